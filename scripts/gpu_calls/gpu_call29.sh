#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --step-size 0.25 --volume 256 256 256 > gpurun_out/bench_sweep_128_025.json 2> gpurun_out/bench_sweep_128_025.err; echo "sweep_exit=$?"; cut -c1-300 gpurun_out/bench_sweep_128_025.json; grep -E "e2e|single" gpurun_out/bench_sweep_128_025.err | tail -3
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --patch 160 --step-size 0.25 --volume 256 256 256 --batch 8 > gpurun_out/bench_sweep_160_025.json 2> gpurun_out/bench_sweep_160_025.err; echo "sweep_exit=$?"; cut -c1-300 gpurun_out/bench_sweep_160_025.json; grep -E "e2e|single" gpurun_out/bench_sweep_160_025.err | tail -3
