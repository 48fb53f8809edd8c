#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/bringup_conv.py brick_c4_32 brick_c4_64 > gpurun_out/bringup_epi2.log 2>&1; echo "bringup_exit=$?"; tail -6 gpurun_out/bringup_epi2.log
timeout 300 python scripts/diag_case.py large 8 > gpurun_out/diag_epi2.log 2>&1; echo "diag_exit=$?"; grep -E "step   0|back-to-back|run_case" gpurun_out/diag_epi2.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r01j.log 2>&1; echo "pytest_exit=$?"; tail -5 gpurun_out/pytest_gpu_r01j.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench_exit=$?"; cat gpurun_out/bench8.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke4.log 2>&1; echo "smoke_exit=$?"; tail -2 gpurun_out/smoke4.log
