cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
G="python bench.py --workload dp --steps 3 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 60 --csv --log-file gpurun_out/r2g_launches_dp.csv $G > gpurun_out/ncu_dp_launches.log 2>&1
echo "dp launch list rc=$?"
python profiles/dp_kernel_times.py gpurun_out/r2g_launches_dp.csv
