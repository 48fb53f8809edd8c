# round 2, call 8 (profiling): launch list of one forward batch of each model + one tile's gather / head / finalize, and
# ncu --set full captures of every brick-kernel launch of the second (warm) pass (dominant layer, the in-consumer-norm
# layer, the kw-packed first layers)
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build()" || exit 1
python scripts/profile_forward.py > gpurun_out/r02_profile_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r02_profile_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python scripts/profile_forward.py > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu list rc=$?"
# brick launches: 7 per forward batch of model 1 (4 batches), then 5 per batch of model 2: the second batch of each is captured
ncu --set full --clock-control none --import-source on -k regex:conv_brick_kernel -s 7 -c 7 -o gpurun_out/r02_prof_brick_m1 python scripts/profile_forward.py > gpurun_out/r02_ncu_full1.log 2>&1; echo "ncu full m1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_brick_kernel -s 33 -c 5 -o gpurun_out/r02_prof_brick_m2 python scripts/profile_forward.py > gpurun_out/r02_ncu_full2.log 2>&1; echo "ncu full m2 rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
