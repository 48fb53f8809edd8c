#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -s > gpurun_out/pytest_gpu_r01d.log 2>&1; echo "pytest_exit=$?"; grep -E "err|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r01d.log | head -40
timeout 400 python scripts/diag_case.py large 16 > gpurun_out/diag8.log 2>&1; echo "diag_exit=$?"
grep -E "norm|engine|run_case|lanes" gpurun_out/diag8.log | cut -c1-210 | grep -E "step +(0|1|25|26) |engine|run_case|lanes"
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench_exit=$?"; cat gpurun_out/bench5.json; tail -3 gpurun_out/bench5.err
