cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large_batch.py -m gpu -q -x > gpurun_out/r2_pytest31.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest31.log
tail -15 gpurun_out/r2_pytest31.log
for f in 1; do
DQN_B200_LB_FUSE_HEAD=$f timeout 300 python bench.py --workload dp --steps 40 --warmup 5 > gpurun_out/r2_dp31_$f.json 2> gpurun_out/r2_dp31_$f.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_dp31_$f.json').read().strip().splitlines()[-1])
print('fuse=$f', d['value'], d['ms_per_step'], d['roofline']['frac'])
PY
done
