# round 2, call 12: forwards in flight (2 lanes x 8 / 12 / 16) on the N=1 bench; fp32-mode per-layer table; configs[0] in fp32 mode
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s --timeout 600 > gpurun_out/r02_pytest12_sharded.log 2>&1; echo "sharded pytest rc=$?"; tail -4 gpurun_out/r02_pytest12_sharded.log
for b in 16 24 32; do
  timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 --batch $b --no-cpu-baseline --no-incumbent --no-hbm > gpurun_out/r02_bench12_b$b.json 2> gpurun_out/r02_bench12_b$b.err; echo "batch $b rc=$?"; grep "resident\|e2e\|single" gpurun_out/r02_bench12_b$b.err
done
BSG_ACT_DTYPE=fp32 python scripts/diag_layers.py 4 > gpurun_out/r02_layers12_fp32.log 2>&1; grep "back-to-back" gpurun_out/r02_layers12_fp32.log
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -s --timeout 800 -k "config1_full" > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_pytest12.log
# ncu --set full of the stride-2 tile-kernel launches (32->64 @128^3: <32, 4>; 64->128 @128^3 and deeper: <64, 2>)
python scripts/profile_forward.py > gpurun_out/r02_profile_plain12.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r02_profile_plain12.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:conv_tc_kernel<32, 4|conv_tc_kernel<64, 2" -c 7 -o gpurun_out/r02_prof_s2 python scripts/profile_forward.py > gpurun_out/r02_ncu_full12.log 2>&1; echo "ncu full s2 rc=$?"
ls -la gpurun_out/r02_prof_s2.ncu-rep
