#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -q -m gpu -x -s -k "without_sliding" > gpurun_out/pytest_full.log 2>&1; echo "pytest_exit=$?"; grep -E "whole-volume|passed|failed|Error|error" gpurun_out/pytest_full.log | head -20; tail -25 gpurun_out/pytest_full.log
