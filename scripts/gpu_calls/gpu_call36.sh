#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python scripts/sanitize_voxelops.py > gpurun_out/sanitize_plain.log 2>&1; rc=$?; echo "plain_exit=$rc"; tail -3 gpurun_out/sanitize_plain.log
if [ $rc -eq 0 ]; then
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 77 python scripts/sanitize_voxelops.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck_exit=$?"; tail -8 gpurun_out/sanitize_memcheck.log
fi
