#!/bin/bash
# one gpurun call: descriptor probe, postproc tests, bench, ncu launch list, ncu full capture of the top conv layer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 60 build/probe_umma_shift > gpurun_out/probe_umma_shift.log 2>&1; echo "probe_exit=$?"
timeout 300 python -m pytest tests/test_gpu_postproc.py -x -q -m gpu > gpurun_out/pytest_postproc2.log 2>&1; echo "postproc_exit=$?"; tail -3 gpurun_out/pytest_postproc2.log
timeout 400 python bench.py --steps 2 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench_exit=$?"; cat gpurun_out/bench3.json
timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench3b.json 2> gpurun_out/bench3b.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu_list_exit=$?"
timeout 120 python scripts/bringup_conv.py perf_64_32_128 > gpurun_out/bringup_p.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 2 -o gpurun_out/prof_conv_64_32_128 \
    python scripts/bringup_conv.py perf_64_32_128 > gpurun_out/ncu_full.log 2>&1; echo "ncu_full_exit=$?"
ls -la gpurun_out
