cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_train_golden.py tests/test_gpu_population.py tests/test_gpu_episode.py -m gpu -q -x > gpurun_out/r2_pytest20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest20.log
tail -30 gpurun_out/r2_pytest20.log
for k in cta cta_tc; do
  timeout 300 python bench.py --workload population --step-kernel $k --steps 128 --warmup 3 > gpurun_out/r2_pop20_$k.json 2> gpurun_out/r2_pop20_$k.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_pop20_$k.json').read().strip().splitlines()[-1])
print('$k', d['value'], d['ms_per_step'], d.get('fp32'), d['timing'], d.get('param_digest_sum'))
PY
done
