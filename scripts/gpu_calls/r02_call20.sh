# round 2, call 20: TMA tensor stores with whole-line (128-byte) rows — kernel tests, per-layer table with the transposed
# convs staged (BSG_TMA_STORE=3) against the direct stores
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_conv_kernels.py -m gpu -q -s --timeout 300 -k "tma" > gpurun_out/r02_pytest20.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest20.log
for m in 0 3; do
  BSG_TMA_STORE=$m timeout 300 python scripts/diag_layers.py 4 > gpurun_out/r02_layers20_tma$m.log 2>&1; echo "tma_store=$m rc=$?"; grep "back-to-back\|convT2" gpurun_out/r02_layers20_tma$m.log | cut -c1-100
done
