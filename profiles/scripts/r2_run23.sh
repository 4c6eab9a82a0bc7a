cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_train_golden.py tests/test_gpu_population.py tests/test_gpu_episode.py -m gpu -q -x -k "not cluster" > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest23.log
tail -30 gpurun_out/r2_pytest23.log
timeout 300 python bench.py --workload population --step-kernel cta_tc --steps 128 --warmup 3 > gpurun_out/r2_pop23.json 2> gpurun_out/r2_pop23.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_pop23.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['timing'], d.get('param_digest_sum'))
PY
P="python bench.py --workload population --step-kernel cta_tc --steps 128 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:dqn_train_tc -s 2 -c 1 -f -o gpurun_out/r2_prof_tc512c $P > gpurun_out/ncu_tc512c.log 2>&1
echo "ncu rc=$?"
