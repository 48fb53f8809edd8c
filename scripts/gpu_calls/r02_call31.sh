# round 2, call 31: ShardedExchange.close() as a collective that raises last — the two-process test + the in-kernel ordering test
cd "$GRAFT_REPO_ROOT"
timeout 150 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 120 -k "two_ranks or signal_orders" > gpurun_out/r02_pytest31.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest31.log
