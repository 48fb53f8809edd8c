#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python scripts/profile_forward.py > gpurun_out/profile_plain2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r01_final.csv \
    python scripts/profile_forward.py > gpurun_out/ncu_list3.log 2>&1; echo "ncu_list_exit=$?"
timeout 200 python scripts/bringup_conv.py perfd_128_64_128_b8 perfb_32_32_128 > gpurun_out/bringup_p3.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_brick -s 4 -c 1 -o gpurun_out/prof_brick_dominant \
    python scripts/bringup_conv.py perfd_128_64_128_b8 > gpurun_out/ncu_full3.log 2>&1; echo "ncu_full_exit=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_brick -s 4 -c 1 -o gpurun_out/prof_brick_32_32 \
    python scripts/bringup_conv.py perfb_32_32_128 > gpurun_out/ncu_full4.log 2>&1; echo "ncu_full2_exit=$?"
cat gpurun_out/bringup_p3.log | sed 's/box .*| //'; tail -2 gpurun_out/ncu_full3.log
