cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for B in 54 70; do
ncu --set full --clock-control none --import-source on -k regex:dqn_train_tc -s 1 -c 1 -f -o gpurun_out/r2_prof_tc2_b$B python profiles/pop_batch_cost.py cta_tc $B > gpurun_out/ncu_tc2_b$B.log 2>&1
echo "ncu rc=$?"
done
