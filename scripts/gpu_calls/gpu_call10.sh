#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py f16 brick_c32_32 k3_stats convT > gpurun_out/bringup7.log 2>&1; echo "bringup_exit=$?"
cat gpurun_out/bringup7.log
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu_r01c.log 2>&1; echo "pytest_exit=$?"; grep -E "err|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r01c.log | head -40
timeout 400 python scripts/diag_case.py large 8 > gpurun_out/diag7.log 2>&1; echo "diag_exit=$?"
grep -E "norm|engine|run_case|lanes" gpurun_out/diag7.log | cut -c1-210
