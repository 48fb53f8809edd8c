#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python scripts/diag_case.py large 8 > gpurun_out/diag6.log 2>&1; echo "diag_exit=$?"
grep -E "run_case|lanes|Error|error" gpurun_out/diag6.log | cut -c1-150; tail -3 gpurun_out/diag6.log
