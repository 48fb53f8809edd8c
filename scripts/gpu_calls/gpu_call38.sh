#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/bringup_conv.py > gpurun_out/bringup_epi1.log 2>&1; echo "bringup_exit=$?"; grep -E "FAIL|failures|perfT|perfs2" gpurun_out/bringup_epi1.log | sed 's/box .*| //'
timeout 300 python scripts/diag_case.py large 8 > gpurun_out/diag_epi1.log 2>&1; echo "diag_exit=$?"; grep -E "convT|back-to-back|step  (1[3-9]|[5-9]) " gpurun_out/diag_epi1.log | cut -c1-110
