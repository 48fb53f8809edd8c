#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py k3_ convT brick_c32 brick_slices brick_c32_24 f16 k1 > gpurun_out/bringup11.log 2>&1; echo "bringup_exit=$?"; grep -c PASS gpurun_out/bringup11.log; grep -E "FAIL|ERROR|failures" gpurun_out/bringup11.log | cut -c1-200
timeout 400 python scripts/diag_case.py large 16 > gpurun_out/diag11.log 2>&1; echo "diag_exit=$?"
grep -E "convT|engine|run_case|lanes" gpurun_out/diag11.log | cut -c1-140
