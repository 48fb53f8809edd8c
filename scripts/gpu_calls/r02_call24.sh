# round 2, call 24: tile-kernel M blocking — kernel tests, per-layer table with the planner's choice (stride 2), everywhere, off
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_conv_kernels.py -m gpu -q -s --timeout 200 -k "m_blocking" > gpurun_out/r02_pytest24.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest24.log
for m in 2 0 1; do
  BSG_MBLOCK=$m timeout 300 python scripts/diag_layers.py 4 > gpurun_out/r02_layers24_mb$m.log 2>&1; echo "mblock=$m rc=$?"; grep "back-to-back" gpurun_out/r02_layers24_mb$m.log | cut -c1-100
done
