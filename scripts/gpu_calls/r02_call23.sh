# round 2, call 23: ncu launch list of the final build (one forward batch of each model + one tile's gather / head / finalize)
# and one --set full capture of the transposed convs with the staged (whole-line) stores
cd "$GRAFT_REPO_ROOT"
python scripts/profile_forward.py > gpurun_out/r02_profile_plain23.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r02_profile_plain23.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches23.csv python scripts/profile_forward.py > gpurun_out/r02_ncu_list23.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel<\(int\)64, \(int\)0, \(int\)2" -s 10 -c 5 -o gpurun_out/r02_prof_convT python scripts/profile_forward.py > gpurun_out/r02_ncu_full23.log 2>&1; echo "ncu full convT rc=$?"
ls -la gpurun_out/r02_prof_convT.ncu-rep
