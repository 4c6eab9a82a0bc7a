cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_train_golden.py tests/test_gpu_population.py tests/test_gpu_episode.py -m gpu -q -x -k "not cluster" > gpurun_out/r2_pytest46.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest46.log
tail -4 gpurun_out/r2_pytest46.log
python profiles/pop_batch_cost.py cta_tc 64 70 80 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:dqn_train_tc -s 1 -c 1 -f -o gpurun_out/r2_prof_tc4_b70 python profiles/pop_batch_cost.py cta_tc 70 > gpurun_out/ncu_tc4_b70.log 2>&1
echo "ncu rc=$?"
