# round 2, call 22: where the 10 % of a case outside the conv stacks goes — blobby post-processing labels, one stream lane
cd "$GRAFT_REPO_ROOT"
for cfg in "--post-labels blobby" "--lanes 1 --batch 8" "--lanes 1 --batch 16"; do
  tag=$(echo $cfg | tr -d ' -')
  timeout 600 python bench.py --gpus 1 --steps 6 --warmup 3 $cfg --no-cpu-baseline --no-incumbent --no-hbm > gpurun_out/r02_bench22_$tag.json 2> gpurun_out/r02_bench22_$tag.err; echo "$cfg rc=$?"; grep "resident\|e2e\|single" gpurun_out/r02_bench22_$tag.err
done
