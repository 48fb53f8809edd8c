# round 2, call 16: whole GPU suite + the driver's N=1 bench command line + CPU arm on the current build
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke16.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke16.log
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 1200 > gpurun_out/r02_pytest16.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest16.log
python scripts/diag_layers.py 4 > gpurun_out/r02_layers16.log 2>&1; grep "back-to-back" gpurun_out/r02_layers16.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench16.json 2> gpurun_out/r02_bench16.err; echo "bench rc=$?"; grep "resident\|e2e\|single\|incumbent\|cpu baseline" gpurun_out/r02_bench16.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_ref16.json 2> gpurun_out/r02_ref16.err; echo "ref rc=$?"; tail -3 gpurun_out/r02_ref16.err
