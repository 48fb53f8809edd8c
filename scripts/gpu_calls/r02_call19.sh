# round 2, call 19: brick epilogue with packed fp32 pairs (FADD2 / FFMA2) and the statistics transpose-reduce once per batch
# item: conv-kernel + network parity tests, per-layer table
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_conv_kernels.py tests/test_gpu_unet.py -m gpu -q --timeout 600 -k "not config1_full and not mirror_equivariance" > gpurun_out/r02_pytest19.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest19.log
timeout 300 python scripts/diag_layers.py 4 > gpurun_out/r02_layers19.log 2>&1; grep "back-to-back\|step   0 \|step  25 \|step   1 " gpurun_out/r02_layers19.log | cut -c1-110
