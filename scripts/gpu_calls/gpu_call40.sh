#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python scripts/profile_forward.py > gpurun_out/profile_plain4.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r01_final3.csv \
    python scripts/profile_forward.py > gpurun_out/ncu_list5.log 2>&1; echo "ncu_list_exit=$?"
