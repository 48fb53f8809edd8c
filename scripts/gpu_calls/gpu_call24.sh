#!/bin/bash
# 8-GPU weak-scaling check of the benchmark (one rank per GPU, cases sharded, no data-path collective)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench_n8_exit=$?"
cut -c1-330 gpurun_out/bench_n8.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n8.json; grep -o '"clocks": {[^}]*}' gpurun_out/bench_n8.json
tail -5 gpurun_out/bench_n8.err
