cd $GRAFT_REPO_ROOT
python profiles/e2e_host_probe.py 8000 2>&1 | tail -5
python profiles/e2e_host_probe.py 8000 2>&1 | tail -5
