# round 2, call 1: build + smoke, conv-kernel tests first (new in-consumer norm transform; a failure there switches the
# rest of the call to BSG_FUSE_NORM=0), the whole GPU test suite, N=1 bench, CPU arm
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke1.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke1.log
timeout 600 python -m pytest tests/test_gpu_conv_kernels.py -q -s --timeout 200 > gpurun_out/r02_pytest1_conv.log 2>&1; rc=$?; echo "conv-kernel tests rc=$rc"; tail -25 gpurun_out/r02_pytest1_conv.log
if [ $rc -ne 0 ]; then export BSG_FUSE_NORM=0; echo "!! continuing with BSG_FUSE_NORM=0"; fi
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 1200 --deselect tests/test_gpu_conv_kernels.py > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r02_bench1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_ref1.json 2> gpurun_out/r02_ref1.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r02_ref1.err
