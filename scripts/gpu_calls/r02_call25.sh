# round 2, call 25: conv-kernel + network parity tests and the N=1 bench (short) with M blocking by the planner's rule
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_conv_kernels.py tests/test_gpu_unet.py tests/test_gpu_config2.py -m gpu -q --timeout 800 -k "not config1_full and not mirror_equivariance and not whole_case" > gpurun_out/r02_pytest25.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest25.log
for m in 0 2 0 2; do
  BSG_MBLOCK=$m timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --no-incumbent --no-hbm > gpurun_out/r02_bench25_mb$m.json 2> gpurun_out/r02_bench25_mb$m.err; echo "mblock=$m rc=$?"; grep "resident\|e2e" gpurun_out/r02_bench25_mb$m.err
done
