cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for v in main c37 spread; do
  if [ $v = main ]; then unset DQN_B200_LIB; else export DQN_B200_LIB=$GRAFT_REPO_ROOT/deep-q-learning_b200/csrc/variants/libdqn_$v.so; fi
  echo "== $v"
  python profiles/pop_batch_cost.py cta_tc 54 64 70 2>&1 | tail -3
done
done
