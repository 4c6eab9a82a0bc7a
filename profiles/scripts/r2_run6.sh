cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -12 gpurun_out/r2_pytest6.log
python bench.py --workload population > gpurun_out/r2_pop1_tail16.json 2> gpurun_out/r2_run6.err
python bench.py --workload population --agents 128 > gpurun_out/r2_pop128_tail16.json 2>> gpurun_out/r2_run6.err
python bench.py --workload single --steps 20000 --warmup 100 --no-cpu-baseline > gpurun_out/r2_single_t16.json 2>> gpurun_out/r2_run6.err
python bench.py --workload single --steps 20000 --warmup 100 --no-cpu-baseline --step-kernel cta > gpurun_out/r2_single_cta.json 2>> gpurun_out/r2_run6.err
python - <<'PY'
import json
for f in ('r2_pop1_tail16','r2_pop128_tail16','r2_single_t16','r2_single_cta'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['roofline'].get('fp32'))
    except Exception as e: print(f,'ERR',e)
PY
tail -5 gpurun_out/r2_run6.err
