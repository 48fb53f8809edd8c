#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 build/probe_tmem_a > gpurun_out/probe_tmem_a.log 2>&1; echo "probe_exit=$?"
timeout 400 python scripts/bringup_conv.py brick convT k3_ perfb perft > gpurun_out/bringup4.log 2>&1; echo "bringup_exit=$?"
tail -45 gpurun_out/bringup4.log
cat gpurun_out/probe_tmem_a.log
