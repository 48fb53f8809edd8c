#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 build/probe_tmem_a > gpurun_out/probe_tmem_a2.log 2>&1; echo "probe_exit=$?"
grep -E "unrolled|correct" gpurun_out/probe_tmem_a2.log
timeout 300 python scripts/diag_case.py large 8 > gpurun_out/diag2.log 2>&1; echo "diag_exit=$?"
grep -E "step|engine|model|run_case|ccl|detect|analyze" gpurun_out/diag2.log | cut -c1-200
