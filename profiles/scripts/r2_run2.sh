cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_large_batch.py::test_real_layer_shapes_consecutive_steps > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -30 gpurun_out/r2_pytest2.log
timeout 300 python bench.py --workload single --steps 20000 --warmup 100 --no-cpu-baseline > gpurun_out/r2_single_push.json 2> gpurun_out/r2_run2.err
timeout 300 python bench.py --workload single --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_single_push20.json 2>> gpurun_out/r2_run2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_single_push*.json')):
    try:
        d=json.load(open(f)); print(f, d['value'], d['e2e']['value'], d['timing'])
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 python profiles/scripts/lb_error_growth.py > gpurun_out/r2_lb_error_growth.log 2>&1
tail -5 gpurun_out/r2_run2.err
