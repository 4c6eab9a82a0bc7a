cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
G="python bench.py --workload dp --batch 8192 --steps 6 --warmup 3"
$G > gpurun_out/r2_dp8192.json 2> gpurun_out/r2_dp8192.err; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_dp8192.json').read().strip().splitlines()[-1]); print('B=8192 on one GPU: ms/step', d['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 80 --csv --log-file gpurun_out/r2_launches_dp8192.csv $G > gpurun_out/ncu_dp8192.log 2>&1
echo "dp launch list rc=$?"
