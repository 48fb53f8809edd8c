#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -x -s -k "config1" > gpurun_out/pytest_config1.log 2>&1; echo "pytest_exit=$?"; grep -E "prob max err|passed|failed|Error|agreement" gpurun_out/pytest_config1.log | head -20
