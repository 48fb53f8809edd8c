# round 2, call 29: head kernel instantiated for 3 classes (the BraTS heads): parity tests + HBM fractions
cd "$GRAFT_REPO_ROOT"
timeout 500 python -m pytest tests/test_gpu_unet.py -m gpu -q --timeout 300 -k "not config1_full and not mirror_equivariance and not brats_architecture" > gpurun_out/r02_pytest29.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest29.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-incumbent > gpurun_out/r02_bench29.json 2> gpurun_out/r02_bench29.err; echo "bench rc=$?"; grep "resident\|e2e" gpurun_out/r02_bench29.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench29.json') if l.startswith('{')][-1])
for h in d['roofline_hbm'][:3]: print(h['kernel'][:50], round(h['ms'],4), round(h['frac'],3))
print(d['result_check']['pass'], d['result_check']['label_agreement_final'])
PY
