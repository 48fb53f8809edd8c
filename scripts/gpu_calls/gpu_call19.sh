#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu_r01h.log 2>&1; echo "pytest_exit=$?"; grep -E "err|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r01h.log | head -30
timeout 400 python bench.py --steps 4 --warmup 3 > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "bench_exit=$?"; cat gpurun_out/bench9.json; grep -E "resident|e2e|single" gpurun_out/bench9.err
