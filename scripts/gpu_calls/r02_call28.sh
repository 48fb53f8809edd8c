# round 2, call 28: finalize kernel with a compile-time class bound (62 registers, 4 blocks per SM): parity tests + its HBM fraction
cd "$GRAFT_REPO_ROOT"
timeout 500 python -m pytest tests/test_gpu_unet.py tests/test_gpu_sharded.py -m gpu -q --timeout 300 -k "not config1_full and not mirror_equivariance and not brats_architecture and not gives_up" > gpurun_out/r02_pytest28.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest28.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-incumbent > gpurun_out/r02_bench28.json 2> gpurun_out/r02_bench28.err; echo "bench rc=$?"; grep "resident\|e2e" gpurun_out/r02_bench28.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench28.json') if l.startswith('{')][-1])
for h in d['roofline_hbm']: print(h['kernel'][:50], round(h['ms'],4), round(h['frac'],3))
print(d['result_check']['pass'], d['result_check']['label_agreement_final'])
PY
