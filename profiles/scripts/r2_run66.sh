cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large_batch.py -m gpu -q -x 2>&1 | tail -1
for v in main prev; do
  if [ $v = main ]; then unset DQN_B200_LIB; else export DQN_B200_LIB=$GRAFT_REPO_ROOT/deep-q-learning_b200/csrc/variants/libdqn_$v.so; fi
  for BB in 65536 8192; do
  G="python bench.py --workload dp --batch $BB --steps 3 --warmup 3"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lb_dh2 -s 3 -c 3 --csv --log-file gpurun_out/r2_dh2.csv $G > /dev/null 2>&1
  echo "$v B=$BB dh2 ns:" $(grep -o 'lb_dh2[^"]*".*' gpurun_out/r2_dh2.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')
  done
done
