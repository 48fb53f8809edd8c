# round 2, call 27: last check of the committed build — smoke, the conv-kernel and sharded test files, a short bench
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke27.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke27.log
timeout 600 python -m pytest tests/test_gpu_conv_kernels.py tests/test_gpu_sharded.py tests/test_gpu_postproc.py -m gpu -q --timeout 300 > gpurun_out/r02_pytest27.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest27.log
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-incumbent > gpurun_out/r02_bench27.json 2> gpurun_out/r02_bench27.err; echo "bench rc=$?"; grep "resident\|e2e" gpurun_out/r02_bench27.err
