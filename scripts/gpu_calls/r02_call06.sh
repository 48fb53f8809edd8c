# round 2, call 6: column-split epilogue groups (CC16 / NT64 / statistics first layer): UNet tests + per-layer table;
# bench A/B: lane streams at high priority (default) vs equal priority, and post-processing on blobby labels
cd "$GRAFT_REPO_ROOT"
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests/test_gpu_conv_kernels.py tests/test_gpu_unet.py -q -s --timeout 600 > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest6.log
python scripts/diag_layers.py 4 > gpurun_out/r02_layers6_default.log 2>&1; echo "diag rc=$?"; grep "back-to-back\|sum of steps\|step   0 " gpurun_out/r02_layers6_default.log
for cfg in "prio:-1:inference" "prio:0:inference" "prio:-1:blobby"; do
  IFS=: read _ prio labels <<< "$cfg"
  BSG_LANE_PRIORITY=$prio timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-incumbent --no-hbm --post-labels $labels > gpurun_out/r02_bench6_p${prio}_${labels}.json 2> gpurun_out/r02_bench6_p${prio}_${labels}.err; echo "bench prio=$prio labels=$labels rc=$?"; grep "resident\|e2e:" gpurun_out/r02_bench6_p${prio}_${labels}.err
done
