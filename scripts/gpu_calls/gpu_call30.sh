#!/bin/bash
# 8-GPU box: latency mode at N=8 (one case sharded over the ranks + NCCL all-reduce) and throughput mode at N=4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 3 --warmup 3 --mode latency --no-cpu-baseline > gpurun_out/bench_n8_latency.json 2> gpurun_out/bench_n8_latency.err; echo "n8_latency_exit=$?"
cut -c1-330 gpurun_out/bench_n8_latency.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n8_latency.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "n4_exit=$?"
cut -c1-330 gpurun_out/bench_n4.json
tail -3 gpurun_out/bench_n8_latency.err
