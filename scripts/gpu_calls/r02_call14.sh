# round 2, call 14: in-kernel rank ordering of the peer exchange (bounded waits) as a single-process three-stream test;
# TMA tensor-store epilogue — kernel tests, then the per-layer table with the stores off / planner's choice / on everywhere;
# ncu --set full of the stride-2 tile-kernel launches
cd "$GRAFT_REPO_ROOT"
timeout 120 python -X faulthandler -m pytest tests/test_gpu_sharded.py -m gpu -q -s --timeout 100 -k "signal" > gpurun_out/r02_pytest14_signal.log 2>&1; echo "signal pytest rc=$?"; tail -12 gpurun_out/r02_pytest14_signal.log
timeout 900 python -m pytest tests/test_gpu_conv_kernels.py -m gpu -q -s --timeout 300 -k "tma" > gpurun_out/r02_pytest14.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest14.log
for m in 2 -1 1; do
  BSG_TMA_STORE=$m timeout 300 python scripts/diag_layers.py 4 > gpurun_out/r02_layers14_tma$m.log 2>&1; echo "tma_store=$m rc=$?"; grep "back-to-back\|convT2" gpurun_out/r02_layers14_tma$m.log | cut -c1-100
done
timeout 300 python scripts/profile_forward.py > gpurun_out/r02_profile_plain14.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r02_profile_plain14.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel<32, 4|conv_tc_kernel<64, 2" -c 7 -o gpurun_out/r02_prof_s2 python scripts/profile_forward.py > gpurun_out/r02_ncu_full14.log 2>&1; echo "ncu full s2 rc=$?"
ls -la gpurun_out/r02_prof_s2.ncu-rep
