cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_train_golden.py tests/test_gpu_population.py tests/test_gpu_episode.py -m gpu -q -x -k "not cluster" > gpurun_out/r2_pytest42.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest42.log
tail -4 gpurun_out/r2_pytest42.log
timeout 300 python bench.py --workload population --step-kernel cta_tc --steps 128 --warmup 3 > gpurun_out/r2_pop42.json 2> gpurun_out/r2_pop42.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_pop42.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
PY
python profiles/pop_batch_cost.py cta_tc 38 64 65 70 80 96 2>&1 | tail -6
