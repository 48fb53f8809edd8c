# round 2, call 32 (2 GPUs): last multi-GPU check of bench.py (configs[2] alone, few steps)
cd "$GRAFT_REPO_ROOT"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --mode latency --steps 4 --warmup 2 --route peer --no-incumbent --no-cpu-baseline --no-hbm > gpurun_out/r02_bench32_n2.json 2> gpurun_out/r02_bench32_n2.err; echo "bench rc=$?"; grep "latency mode\|Error\|error" gpurun_out/r02_bench32_n2.err | head -4; tail -c 300 gpurun_out/r02_bench32_n2.json
