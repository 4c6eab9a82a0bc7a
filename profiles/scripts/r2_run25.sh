cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python profiles/host_breakdown.py 2>&1 | tail -8
