cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2_bench_noflags.json 2> gpurun_out/r2_bench_noflags.err ) 2>&1 | grep real; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_noflags.json').read().strip().splitlines()[-1])
print('steps', d['steps'], 'warmup', d['warmup'], 'value', d['value'], 'e2e', d['e2e']['value'], 'timing', d.get('timing'), 'launches', d['gpu_launches'])
print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d['extras'].items()})
PY
