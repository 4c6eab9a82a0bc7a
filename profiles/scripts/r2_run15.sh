cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest15.log
tail -5 gpurun_out/r2_pytest15.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench15.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'timing', d.get('timing'), 'cpu', d['cpu_baseline'])
for k,v in d['extras'].items():
    if isinstance(v, dict): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), (v.get('roofline') or {}).get('frac'))
    else: print(k, v)
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench15_ref.json 2>> gpurun_out/r2_bench15.err ) 2>&1 | grep real
cat gpurun_out/r2_bench15_ref.json | cut -c1-600
