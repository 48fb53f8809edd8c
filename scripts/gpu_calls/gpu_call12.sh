#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py k3_ convT brick_c32 brick_c64 brick_stats f16_tile k1 > gpurun_out/bringup8.log 2>&1; echo "bringup_exit=$?"; grep -c PASS gpurun_out/bringup8.log; grep -E "FAIL|ERROR|failures" gpurun_out/bringup8.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r01e.log 2>&1; echo "pytest_exit=$?"; tail -3 gpurun_out/pytest_gpu_r01e.log
timeout 400 python scripts/diag_case.py large 16 > gpurun_out/diag9.log 2>&1; echo "diag_exit=$?"
grep -E "step|engine|run_case|lanes|analyze" gpurun_out/diag9.log | cut -c1-215
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench_exit=$?"; cat gpurun_out/bench6.json
