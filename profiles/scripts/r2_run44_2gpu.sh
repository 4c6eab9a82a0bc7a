cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi_process.py tests/test_gpu_large_batch.py -m gpu -q -x > gpurun_out/r2h_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_2gpu.log; tail -3 gpurun_out/r2h_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2_run44.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench_n2.json').read().strip().splitlines()[-1])
print('n2 value', d['value'], 'e2e', d['e2e']['value'], 'timing', d['timing'])
for k,v in d['extras'].items():
    if isinstance(v, dict): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'), v.get('replicas'))
    else: print(k, v)
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_run44.err | tail -8
