#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r01b.log 2>&1; echo "pytest_exit=$?"; tail -15 gpurun_out/pytest_gpu_r01b.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench_exit=$?"; cat gpurun_out/bench4.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke_exit=$?"; tail -2 gpurun_out/smoke2.log
