#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python scripts/profile_forward.py > gpurun_out/profile_plain3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r01_final2.csv \
    python scripts/profile_forward.py > gpurun_out/ncu_list4.log 2>&1; echo "ncu_list_exit=$?"
timeout 200 python scripts/bringup_conv.py perfs2_32_64_128_b8 > gpurun_out/bringup_p5.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 4 -c 1 -o gpurun_out/prof_s2_unrolled \
    python scripts/bringup_conv.py perfs2_32_64_128_b8 > gpurun_out/ncu_full7.log 2>&1; echo "ncu_full_exit=$?"
cat gpurun_out/bringup_p5.log | sed 's/box .*| //' | tail -3
