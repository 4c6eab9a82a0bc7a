cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P="python bench.py --workload population --step-kernel cta_tc --steps 128 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:dqn_train_tc -s 2 -c 1 -f -o gpurun_out/r2_prof_tc512 $P > gpurun_out/ncu_tc512.log 2>&1
echo "ncu rc=$?"
