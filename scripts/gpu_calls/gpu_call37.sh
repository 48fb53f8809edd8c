#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python scripts/bringup_conv.py perfT_64_64_64 > gpurun_out/bringup_p6.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o gpurun_out/prof_convT2 \
    python scripts/bringup_conv.py perfT_64_64_64 > gpurun_out/ncu_full8.log 2>&1; echo "ncu_full_exit=$?"
cat gpurun_out/bringup_p6.log | sed 's/box .*| //' | tail -3
