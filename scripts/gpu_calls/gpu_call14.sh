#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python scripts/profile_forward.py > gpurun_out/profile_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r01_brick.csv \
    python scripts/profile_forward.py > gpurun_out/ncu_list2.log 2>&1; echo "ncu_list_exit=$?"
timeout 120 python scripts/bringup_conv.py perfh_128_64_64_f16_stats perfh_64_32_128_bf16 > gpurun_out/bringup_p2.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_brick -c 2 -o gpurun_out/prof_brick \
    python scripts/bringup_conv.py perfh_128_64_64_f16_stats perfh_64_32_128_bf16 > gpurun_out/ncu_full2.log 2>&1; echo "ncu_full_exit=$?"
tail -3 gpurun_out/profile_plain.log; tail -3 gpurun_out/ncu_list2.log; tail -4 gpurun_out/ncu_full2.log; ls -la gpurun_out | tail -6
