# round 2, call 17: three stream lanes (3 x 8 forwards in flight) against the default two
cd "$GRAFT_REPO_ROOT"
for cfg in "16 2" "24 3" "18 3"; do
  set -- $cfg
  timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 --batch $1 --lanes $2 --no-cpu-baseline --no-incumbent --no-hbm > gpurun_out/r02_bench17_b$1_l$2.json 2> gpurun_out/r02_bench17_b$1_l$2.err; echo "batch $1 lanes $2 rc=$?"; grep "resident\|e2e\|single" gpurun_out/r02_bench17_b$1_l$2.err
done
