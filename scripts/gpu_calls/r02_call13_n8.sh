# round 2, call 13 (8 GPUs): configs[2] alone — one case sharded over 8 GPUs, both exchange routes
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --mode latency --steps 20 --warmup 3 --no-incumbent --no-cpu-baseline --no-hbm > gpurun_out/r02_bench13_n8.json 2> gpurun_out/r02_bench13_n8.err; echo "bench rc=$?"; grep "latency mode" gpurun_out/r02_bench13_n8.err | sort | uniq | head; grep -i "error\|Traceback" gpurun_out/r02_bench13_n8.err | head -5
