#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r01f.log 2>&1; echo "pytest_exit=$?"; tail -5 gpurun_out/pytest_gpu_r01f.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench_exit=$?"; cat gpurun_out/bench7.json
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; echo "smoke_exit=$?"; tail -2 gpurun_out/smoke3.log
