cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for k in cta cta_tc; do
  timeout 300 python bench.py --workload population --step-kernel $k --steps 128 --warmup 3 > gpurun_out/r2_pop18_$k.json 2> gpurun_out/r2_pop18_$k.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_pop18_$k.json').read().strip().splitlines()[-1])
print('$k', d['value'], d['ms_per_step'], d.get('fp32'), d['timing'], d.get('param_digest_sum'))
PY
done
