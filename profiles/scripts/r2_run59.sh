cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_train_golden.py tests/test_gpu_session.py tests/test_gpu_episode.py tests/test_gpu_training_loop.py -m gpu -q -x 2>&1 | tail -2
for v in main prev; do
  if [ $v = main ]; then unset DQN_B200_LIB; else export DQN_B200_LIB=$GRAFT_REPO_ROOT/deep-q-learning_b200/csrc/variants/libdqn_$v.so; fi
  echo "== $v"
  for K in 20 20000; do
  timeout 300 python bench.py --workload single --steps $K --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps', d['steps'], 'value', d['value'], 'e2e', d['e2e']['value'])"
  done
done
