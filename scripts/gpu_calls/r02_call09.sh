# round 2, call 9: whole GPU suite + the driver's N=1 bench command line + CPU arm on the current build
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke9.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke9.log
timeout 2400 python -m pytest tests -m gpu -q -s --timeout 1200 > gpurun_out/r02_pytest9.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest9.log
python scripts/diag_layers.py 4 > gpurun_out/r02_layers9.log 2>&1; grep "back-to-back\|step  26 " gpurun_out/r02_layers9.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err; echo "bench rc=$?"; grep "resident\|e2e\|single\|incumbent\|cpu baseline" gpurun_out/r02_bench9.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_ref9.json 2> gpurun_out/r02_ref9.err; echo "ref rc=$?"; tail -3 gpurun_out/r02_ref9.err
