cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --workload population --step-kernel cta_tc --steps 128 --warmup 3 > gpurun_out/r2_pop41.json 2> gpurun_out/r2_pop41.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_pop41.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
PY
for B in 54; do
ncu --set full --clock-control none --import-source on -k regex:dqn_train_tc -s 1 -c 1 -f -o gpurun_out/r2_prof_tc3_b$B python profiles/pop_batch_cost.py cta_tc $B > gpurun_out/ncu_tc3_b$B.log 2>&1
echo "ncu rc=$?"
done
