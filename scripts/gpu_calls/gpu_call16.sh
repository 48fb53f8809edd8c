#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py k3_ convT brick f16 k1 perfb > gpurun_out/bringup12.log 2>&1; echo "bringup_exit=$?"; grep -c PASS gpurun_out/bringup12.log; grep -E "FAIL|ERROR|failures" gpurun_out/bringup12.log | cut -c1-200; grep perfb gpurun_out/bringup12.log | sed 's/box .*| //' 
timeout 400 python scripts/diag_case.py large 16 > gpurun_out/diag12.log 2>&1; echo "diag_exit=$?"
grep -E "step|engine|run_case|lanes" gpurun_out/diag12.log | cut -c1-100
