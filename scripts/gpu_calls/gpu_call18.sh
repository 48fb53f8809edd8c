#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py k3_8cube k3_4cube k3_s2_to4 perf_320 convT_big > gpurun_out/bringup13.log 2>&1; echo "bringup_exit=$?"; grep -E "PASS|FAIL|ERROR|failures" gpurun_out/bringup13.log | cut -c1-190
timeout 600 python -m pytest tests/test_gpu_unet.py -q -m gpu > gpurun_out/pytest_gpu_r01g.log 2>&1; echo "pytest_exit=$?"; tail -3 gpurun_out/pytest_gpu_r01g.log
timeout 400 python scripts/diag_case.py large 16 > gpurun_out/diag13.log 2>&1; echo "diag_exit=$?"
grep -E "step +(8|9|10|11|13|14) |engine|run_case|lanes" gpurun_out/diag13.log | cut -c1-175
