#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/bench_hbm_kernels.py > gpurun_out/hbm_kernels2.log 2>&1; echo "exit=$?"; cat gpurun_out/hbm_kernels2.log | tail -32
