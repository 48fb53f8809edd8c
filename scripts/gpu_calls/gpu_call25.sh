#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/bringup_conv.py > gpurun_out/bringup_modes.log 2>&1; echo "bringup_exit=$?"; grep -E "FAIL|failures|perfs2|perfp_s2" gpurun_out/bringup_modes.log | sed 's/box .*| //'
timeout 300 python scripts/diag_case.py large 8 > gpurun_out/diag_modes.log 2>&1; echo "diag_exit=$?"; grep -E "step  |back-to-back|run_case" gpurun_out/diag_modes.log | cut -c1-120
