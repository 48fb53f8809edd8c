# round 2, call 3: tile-epilogue fix (register array no longer address-taken), in-consumer norm for kw-fused plans only,
# kw-packed first layer (9 taps instead of 27): conv-kernel tests, the UNet / configs[1] / sharded tests, per-layer table, bench
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke3.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke3.log
timeout 600 python -m pytest tests/test_gpu_conv_kernels.py -q -s --timeout 200 > gpurun_out/r02_pytest3_conv.log 2>&1; rc=$?; echo "conv-kernel tests rc=$rc"; tail -6 gpurun_out/r02_pytest3_conv.log
python scripts/diag_layers.py 4 > gpurun_out/r02_layers3_default.log 2>&1; echo "diag rc=$?"; grep "back-to-back\|sum of steps" gpurun_out/r02_layers3_default.log
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 1200 --deselect tests/test_gpu_conv_kernels.py > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest3.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"; grep "resident\|e2e\|single\|incumbent" gpurun_out/r02_bench3.err
