cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P="python bench.py --workload population --step-kernel cta_tc --steps 128 --warmup 3"
$P > gpurun_out/r2h_pop_plain.json 2> gpurun_out/r2h_pop_plain.err && ncu --set full --clock-control none --import-source on -k regex:dqn_train_tc -s 2 -c 1 -f -o gpurun_out/r2h_prof_pop $P > gpurun_out/ncu_pop48.log 2>&1
echo "ncu pop rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2h_launches_pop.csv $P > gpurun_out/ncu_pop48b.log 2>&1
echo "launch list rc=$?"
