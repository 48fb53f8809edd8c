#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py stats k3_4cube brick > gpurun_out/bringup9.log 2>&1; echo "bringup_exit=$?"; grep -c PASS gpurun_out/bringup9.log; grep -E "FAIL|ERROR|failures|stats" gpurun_out/bringup9.log | cut -c1-200
timeout 400 python scripts/diag_case.py large 16 > gpurun_out/diag10.log 2>&1; echo "diag_exit=$?"
grep -E "norm|engine|run_case|lanes|analyze" gpurun_out/diag10.log | cut -c1-215 | grep -E "step +(0|1|23|25|26) |engine|run_case|lanes|analyze"
