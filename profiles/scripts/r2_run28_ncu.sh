cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --gpus 1 --steps 20 --warmup 5"
$B > gpurun_out/r2f_bench_plain.json 2> gpurun_out/r2f_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2f_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
S="python bench.py --gpus 1 --steps 500 --warmup 3 --profile-only"
$S > gpurun_out/plain_single.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dqn_train_cluster -s 3 -c 1 -f -o gpurun_out/r2f_prof_cluster $S > gpurun_out/ncu_cluster.log 2>&1
echo "ncu cluster rc=$?"
G="python bench.py --workload dp --steps 3 --warmup 3"
$G > gpurun_out/plain_dp.log 2>&1 && ncu --set full --clock-control none -k regex:gemm_tc -s 16 -c 4 -f -o gpurun_out/r2f_prof_gemm $G > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 60 --csv --log-file gpurun_out/r2f_launches_dp.csv $G > gpurun_out/ncu_dp_launches.log 2>&1
echo "dp launch list rc=$?"
R="python bench.py --workload replay --steps 20 --warmup 3"
$R > gpurun_out/plain_replay.log 2>&1 && ncu --set full --clock-control none -k regex:replay_ -s 6 -c 2 -f -o gpurun_out/r2f_prof_replay $R > gpurun_out/ncu_replay.log 2>&1
echo "ncu replay rc=$?"
ls -la gpurun_out | tail -14
