#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -s -k "sweep or folds" > gpurun_out/pytest_gpu_r01i.log 2>&1; echo "pytest_exit=$?"; grep -E "err|passed|failed|FAILED|Error|decisive" gpurun_out/pytest_gpu_r01i.log | head
timeout 500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --patch 160 --step-size 0.5 --volume 256 256 256 --batch 8 > gpurun_out/bench_sweep_160.json 2> gpurun_out/bench_sweep_160.err; echo "sweep_exit=$?"; cut -c1-400 gpurun_out/bench_sweep_160.json; grep -E "resident|e2e|engines" gpurun_out/bench_sweep_160.err
timeout 500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --step-size 0.5 --volume 256 256 256 > gpurun_out/bench_sweep_128.json 2> gpurun_out/bench_sweep_128.err; echo "sweep_exit=$?"; cut -c1-300 gpurun_out/bench_sweep_128.json; grep -E "resident|e2e" gpurun_out/bench_sweep_128.err
