#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py stats k3_4cube k3_c128 brick_c4_64 > gpurun_out/bringup6.log 2>&1; echo "bringup_exit=$?"
cat gpurun_out/bringup6.log
timeout 300 python scripts/diag_case.py large 8 > gpurun_out/diag4.log 2>&1; echo "diag_exit=$?"
grep -E "engine 1|model|run_case|back-to-back" gpurun_out/diag4.log | cut -c1-150
grep -E "norm" gpurun_out/diag4.log | cut -c1-150
