# round 2, call 4 (2 GPUs): the N > 1 bench line — cohort throughput + the configs[2] latency record (peer route and NCCL route)
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"; grep "bench " gpurun_out/r02_bench_n2.err | tail -25; tail -5 gpurun_out/r02_bench_n2.err
