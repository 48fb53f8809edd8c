#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_voxelops.py tests/test_gpu_postproc.py tests/test_gpu_preprocess.py -q -m gpu -x > gpurun_out/pytest_voxelops.log 2>&1; echo "pytest_exit=$?"; tail -40 gpurun_out/pytest_voxelops.log
