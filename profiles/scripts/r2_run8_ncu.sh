cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --gpus 1 --steps 20 --warmup 5"
$B > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
P="python bench.py --workload population --steps 128 --warmup 3"
$P > gpurun_out/plain_pop.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dqn_train_fused -s 2 -c 1 -f -o gpurun_out/r2_prof_fused $P > gpurun_out/ncu_fused.log 2>&1
echo "ncu fused rc=$?"
G="python bench.py --workload dp --steps 3 --warmup 3"
$G > gpurun_out/plain_dp.log 2>&1 && ncu --set full --clock-control none -k regex:gemm_tc -s 16 -c 4 -f -o gpurun_out/r2_prof_gemm $G > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 60 --csv --log-file gpurun_out/r2_launches_dp.csv $G > gpurun_out/ncu_dp_launches.log 2>&1
echo "dp launch list rc=$?"
ls -la gpurun_out | tail -12
