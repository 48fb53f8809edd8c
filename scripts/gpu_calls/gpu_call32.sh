#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_drivers.py tests/test_gpu_cli.py -q -m gpu -x > gpurun_out/pytest_drivers.log 2>&1; echo "pytest_exit=$?"; tail -60 gpurun_out/pytest_drivers.log
