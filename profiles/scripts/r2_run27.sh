cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_session.py -m gpu -q -x 2>&1 | tail -1; done
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest27.log
tail -4 gpurun_out/r2_pytest27.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench27.json 2> gpurun_out/r2_bench27.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench27.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'timing', d.get('timing'))
for k,v in d['extras'].items():
    if isinstance(v, dict): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), (v.get('roofline') or {}).get('frac'))
    else: print(k, v)
PY
