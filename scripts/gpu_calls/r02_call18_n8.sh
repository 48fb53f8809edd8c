# round 2, call 18 (8 GPUs): the driver's N=8 command line — cohort throughput (configs[3]) + the configs[2] latency record
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench18_n8.json 2> gpurun_out/r02_bench18_n8.err; echo "bench rc=$?"; grep "latency mode\|resident\|e2e" gpurun_out/r02_bench18_n8.err | sort | uniq | head -12; grep -i "error\|Traceback" gpurun_out/r02_bench18_n8.err | head -5
