# round 2, call 15: whole sharded test file (in-kernel ordering + two-rank worker), tile kernel tests, per-layer table of the
# default build (epilogue mode is a compile-time parameter again) and with the TMA stores everywhere, ncu of the stride-2 launches
cd "$GRAFT_REPO_ROOT"
timeout 400 python -X faulthandler -u -m pytest tests/test_gpu_sharded.py -m gpu -v -s --timeout 300 > gpurun_out/r02_pytest15_sharded.log 2>&1; echo "sharded pytest rc=$?"; grep "PASSED\|FAILED\|passed\|failed" gpurun_out/r02_pytest15_sharded.log | tail -12
timeout 900 python -m pytest tests/test_gpu_conv_kernels.py -m gpu -q -s --timeout 300 > gpurun_out/r02_pytest15.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest15.log
for m in 0 1; do
  BSG_TMA_STORE=$m timeout 300 python scripts/diag_layers.py 4 > gpurun_out/r02_layers15_tma$m.log 2>&1; echo "tma_store=$m rc=$?"; grep "back-to-back\|convT2" gpurun_out/r02_layers15_tma$m.log | cut -c1-100
done
timeout 300 python scripts/profile_forward.py > gpurun_out/r02_profile_plain15.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r02_profile_plain15.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel<\(int\)32, \(int\)4|conv_tc_kernel<\(int\)64, \(int\)2" -c 7 -o gpurun_out/r02_prof_s2 python scripts/profile_forward.py > gpurun_out/r02_ncu_full15.log 2>&1; echo "ncu full s2 rc=$?"
ls -la gpurun_out/r02_prof_s2.ncu-rep
