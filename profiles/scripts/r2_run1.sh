set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -15 gpurun_out/r2_pytest1.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench1.err
for v in bar0 bar1; do
  DQN_B200_LIB=$GRAFT_REPO_ROOT/deep-q-learning_b200/csrc/variants/libdqn_$v.so timeout 300 python bench.py --workload single --steps 20000 --warmup 100 --no-cpu-baseline > gpurun_out/r2_single_$v.json 2>> gpurun_out/r2_bench1.err
done
timeout 300 python bench.py --workload single --steps 20000 --warmup 100 --no-cpu-baseline > gpurun_out/r2_single_bar2.json 2>> gpurun_out/r2_bench1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_single_*.json')):
    try:
        d=json.load(open(f)); print(f, d['value'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
