cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large_batch.py -m gpu -q -x 2>&1 | tail -2
G="python bench.py --workload dp --batch 8192 --steps 6 --warmup 3"
$G > gpurun_out/r2_dp8192b.json 2> gpurun_out/r2_dp8192b.err; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_dp8192b.json').read().strip().splitlines()[-1]); print('B=8192 on one GPU: ms/step', d['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 80 --csv --log-file gpurun_out/r2_launches_dp8192b.csv $G > gpurun_out/ncu_dp8192b.log 2>&1
echo "dp launch list rc=$?"
python bench.py --workload dp --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=65536: ms/step', d['ms_per_step'])"
