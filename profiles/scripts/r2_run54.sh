cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest54.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest54.log
tail -4 gpurun_out/r2_pytest54.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench54.json 2> gpurun_out/r2_bench54.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench54.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'timing', d.get('timing'))
for k,v in d['extras'].items():
    if isinstance(v, dict): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), (v.get('roofline') or {}).get('frac'))
    else: print(k, v)
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref54.json 2> gpurun_out/r2_ref54.err ) 2>&1 | grep real; tail -c 300 gpurun_out/r2_ref54.json
