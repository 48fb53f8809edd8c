#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py brick perfb > gpurun_out/bringup5.log 2>&1; echo "bringup_exit=$?"
cat gpurun_out/bringup5.log
timeout 300 python scripts/diag_case.py large 8 > gpurun_out/diag3.log 2>&1; echo "diag_exit=$?"
grep -E "step|engine|model|run_case" gpurun_out/diag3.log | cut -c1-150
