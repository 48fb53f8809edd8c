# round 2, call 2: per-layer tables with / without the in-consumer norm transform and the fp16 range guard (model 2 came out
# 17 % slower than round 1 in call 1), re-run of the two failed tests, bench at 32 forwards in flight
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build()" || exit 1
python scripts/diag_layers.py 4 > gpurun_out/r02_layers_default.log 2>&1; echo "diag default rc=$?"
BSG_FUSE_NORM=0 python scripts/diag_layers.py 4 > gpurun_out/r02_layers_nofuse.log 2>&1; echo "diag nofuse rc=$?"
BSG_FUSE_NORM=0 BSG_OVERFLOW_GUARD=0 python scripts/diag_layers.py 4 > gpurun_out/r02_layers_nofuse_noguard.log 2>&1; echo "diag nofuse noguard rc=$?"
grep "back-to-back\|sum of steps" gpurun_out/r02_layers_*.log
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_config2.py::test_fp16_range_guard_reruns_in_bf16 tests/test_gpu_conv_kernels.py -q -s --timeout 600 > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest2.log
BSG_FUSE_NORM=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-incumbent --no-hbm > gpurun_out/r02_bench2_nofuse.json 2> gpurun_out/r02_bench2_nofuse.err; echo "bench nofuse rc=$?"; grep "resident\|e2e" gpurun_out/r02_bench2_nofuse.err
