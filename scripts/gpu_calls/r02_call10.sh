# round 2, call 10: fp32-equivalent mode (fp16x3 split operands) against the fp32 oracle
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests/test_gpu_unet.py -m gpu -q -s --timeout 900 -k "fp32_mode" > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r02_pytest10.log
