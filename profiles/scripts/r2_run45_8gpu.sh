cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n8.json 2> gpurun_out/r2_run45.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench_n8.json').read().strip().splitlines()[-1])
print('n8 value', d['value'], 'e2e', d['e2e']['value'], 'timing', d['timing'])
for k,v in d['extras'].items():
    if isinstance(v, dict): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), v.get('replicas'))
    else: print(k, v)
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_run45.err | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n4.json 2> gpurun_out/r2_run45b.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench_n4.json').read().strip().splitlines()[-1])
print('n4 value', d['value'], 'e2e', d['e2e']['value'])
for k,v in d['extras'].items():
    if isinstance(v, dict): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), v.get('replicas'))
PY
