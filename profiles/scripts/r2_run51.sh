cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_train_golden.py tests/test_gpu_population.py tests/test_gpu_episode.py -m gpu -q -x -k "not cluster" > gpurun_out/r2_pytest51.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest51.log
tail -4 gpurun_out/r2_pytest51.log
for v in main; do
  if [ $v = main ]; then unset DQN_B200_LIB; else export DQN_B200_LIB=$GRAFT_REPO_ROOT/deep-q-learning_b200/csrc/variants/libdqn_$v.so; fi
  echo "== $v"
  python profiles/pop_batch_cost.py cta_tc 54 64 70 80 2>&1 | tail -4
done
unset DQN_B200_LIB
timeout 300 python bench.py --workload population --step-kernel cta_tc --steps 128 --warmup 3 > gpurun_out/r2_pop51.json 2> gpurun_out/r2_pop51.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_pop51.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
PY
