cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large_batch.py -m gpu -q -x 2>&1 | tail -1
for v in main prev main prev; do
  if [ $v = main ]; then unset DQN_B200_LIB; else export DQN_B200_LIB=$GRAFT_REPO_ROOT/deep-q-learning_b200/csrc/variants/libdqn_$v.so; fi
  python bench.py --workload dp --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v B=65536: ms/step', d['ms_per_step'])"
done
unset DQN_B200_LIB
G="python bench.py --workload dp --steps 3 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lb_layer1 -s 3 -c 3 --csv --log-file gpurun_out/r2_layer1.csv $G > /dev/null 2>&1
grep -o 'lb_layer1[^"]*".*' gpurun_out/r2_layer1.csv | awk -F'","' '{print $NF}' | tail -3
