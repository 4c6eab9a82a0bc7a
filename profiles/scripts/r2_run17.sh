cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for k in cta cta_tc; do
  DQN_B200_STEP_KERNEL=$k timeout 300 python bench.py --workload population --steps 128 --warmup 3 > gpurun_out/r2_pop17_$k.json 2> gpurun_out/r2_pop17_$k.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_pop17_$k.json').read().strip().splitlines()[-1])
print('$k', d['value'], d['ms_per_step'], d.get('fp32'), d['timing'], d.get('param_digest_sum'))
PY
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest17.log
tail -5 gpurun_out/r2_pytest17.log
