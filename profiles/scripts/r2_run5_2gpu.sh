cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out profiles
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi_process.py -q -m gpu > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_2gpu.log
tail -8 gpurun_out/r2_pytest_2gpu.log
timeout 600 python bench.py --gpus 2 --workload dp --steps 40 > gpurun_out/r2_dp2_p2p.json 2> gpurun_out/r2_run5.err
timeout 600 python bench.py --gpus 2 --workload dp --steps 40 --collective nccl > gpurun_out/r2_dp2_nccl.json 2>> gpurun_out/r2_run5.err
timeout 600 python bench.py --gpus 2 --workload population > gpurun_out/r2_pop2.json 2>> gpurun_out/r2_run5.err
python - <<'PY'
import json
for f in ('r2_dp2_p2p','r2_dp2_nccl','r2_pop2'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d.get('replicas'))
    except Exception as e: print(f,'ERR',e)
PY
tail -5 gpurun_out/r2_run5.err
