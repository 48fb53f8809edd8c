#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py perfT_64_64_64 perfs2 > gpurun_out/bringup_p4.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o gpurun_out/prof_convT \
    python scripts/bringup_conv.py perfT_64_64_64 > gpurun_out/ncu_full5.log 2>&1; echo "ncu_full_exit=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o gpurun_out/prof_s2 \
    python scripts/bringup_conv.py perfs2 > gpurun_out/ncu_full6.log 2>&1; echo "ncu_full2_exit=$?"
cat gpurun_out/bringup_p4.log | sed 's/box .*| //'
