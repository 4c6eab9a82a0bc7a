cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in main late; do
  if [ $v = main ]; then unset DQN_B200_LIB; else export DQN_B200_LIB=$GRAFT_REPO_ROOT/deep-q-learning_b200/csrc/variants/libdqn_$v.so; fi
  echo "== $v"
  timeout 300 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_population.py -m gpu -q -x -k "not cluster" 2>&1 | tail -1
  python profiles/pop_batch_cost.py cta_tc 54 64 70 2>&1 | tail -3
  timeout 300 python bench.py --workload population --step-kernel cta_tc --steps 128 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'])"
done
