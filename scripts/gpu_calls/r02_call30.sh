# round 2, call 30: the whole GPU suite on the last build of the round
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke30.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke30.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r02_pytest30.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest30.log
