cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_session.py -m gpu -q -x > gpurun_out/r2_pytest26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest26.log
tail -5 gpurun_out/r2_pytest26.log
python profiles/host_breakdown.py 2>&1 | tail -8
