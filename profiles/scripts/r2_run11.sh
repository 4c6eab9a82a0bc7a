cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large_batch.py -m gpu -q > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest11.log
tail -6 gpurun_out/r2_pytest11.log
python bench.py --workload dp --steps 30 > gpurun_out/r2_dp1_cg2.json 2> gpurun_out/r2_run11.err
DQN_B200_GEMM_CG=1 python bench.py --workload dp --steps 30 > gpurun_out/r2_dp1_cg1.json 2>> gpurun_out/r2_run11.err
python - <<'PY'
import json
for f in ('r2_dp1_cg2','r2_dp1_cg1'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['roofline']['frac_of_mode_ceiling'])
PY
tail -3 gpurun_out/r2_run11.err
