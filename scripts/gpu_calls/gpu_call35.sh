#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BSG_ACT_DTYPE=fp16 timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -s -k "config1 or full_patch or sigmoid_all" > gpurun_out/pytest_config1_fp16.log 2>&1; echo "pytest_exit=$?"; grep -E "prob max err|passed|failed|Error|agreement|sigmoid max err" gpurun_out/pytest_config1_fp16.log | head -20
