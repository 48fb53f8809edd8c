# round 2, call 11 (2 GPUs): configs[2] alone (one case sharded) with the balanced batch and two cases submitted ahead
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode latency --steps 10 --warmup 3 > gpurun_out/r02_bench11_n2.json 2> gpurun_out/r02_bench11_n2.err; echo "bench rc=$?"; grep "latency mode\|Error\|error" gpurun_out/r02_bench11_n2.err | head
