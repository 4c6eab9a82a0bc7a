cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_session.py -m gpu -q -x 2>&1 | tail -1; done
timeout 900 python profiles/session_soak.py 2>&1 | tail -4
