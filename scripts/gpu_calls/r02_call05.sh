# round 2, call 5: peer buffers now mapped with the library's own CUDA IPC calls (opened on the consuming device): the
# two-process sharded test; conv-kernel tests and per-layer table after the CC16/NT64 epilogue change
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_conv_kernels.py tests/test_gpu_unet.py -q -s --timeout 600 > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest5.log
python scripts/diag_layers.py 4 > gpurun_out/r02_layers5_default.log 2>&1; echo "diag rc=$?"; grep "back-to-back\|sum of steps\|step   0 \|step  2[56] " gpurun_out/r02_layers5_default.log
