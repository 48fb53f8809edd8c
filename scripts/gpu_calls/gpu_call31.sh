#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/bench_hbm_kernels.py > gpurun_out/hbm_kernels3.log 2>&1; echo "exit=$?"; grep -E "gather|finalize|head_tta" gpurun_out/hbm_kernels3.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r01k.log 2>&1; echo "pytest_exit=$?"; tail -5 gpurun_out/pytest_gpu_r01k.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench_exit=$?"; cut -c1-300 gpurun_out/bench10.json
