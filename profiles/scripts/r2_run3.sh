cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -12 gpurun_out/r2_pytest3.log
P="python bench.py --profile --workload single --steps 500 --warmup 10"
$P > gpurun_out/plain_cluster.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dqn_train_cluster -s 3 -c 1 -f -o gpurun_out/r2_prof_cluster $P > gpurun_out/ncu_cluster.log 2>&1
echo "ncu cluster rc=$?"
R="python bench.py --workload replay --steps 10 --warmup 3"
$R > gpurun_out/plain_replay.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:replay_gather -s 6 -c 1 -f -o gpurun_out/r2_prof_gather $R > gpurun_out/ncu_gather.log 2>&1
echo "ncu gather rc=$?"
ncu --set full --clock-control none --import-source on -k regex:replay_store -s 20 -c 1 -f -o gpurun_out/r2_prof_store $R > gpurun_out/ncu_store.log 2>&1
echo "ncu store rc=$?"
Q="python bench.py --workload per --steps 10 --warmup 3"
$Q > gpurun_out/plain_per.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:per_ -s 40 -c 12 --csv --log-file gpurun_out/r2_per_launches.csv $Q > gpurun_out/ncu_per.log 2>&1
echo "ncu per rc=$?"
ls -la gpurun_out
