cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
tail -12 gpurun_out/r2_pytest4.log
python bench.py --workload replay > gpurun_out/r2_replay_96.json 2> gpurun_out/r2_run4.err
DQN_B200_RECORD_BYTES=128 python bench.py --workload replay > gpurun_out/r2_replay_128.json 2>> gpurun_out/r2_run4.err
python bench.py --workload dp --steps 30 > gpurun_out/r2_dp1.json 2>> gpurun_out/r2_run4.err
python - <<'PY'
import json
for f in ('r2_replay_96','r2_replay_128'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value']/1e9, 'G samples/s', d['roofline']['frac'], 'store', d['replay_stores_per_sec']/1e9, d['roofline']['store']['frac'])
d=json.load(open('gpurun_out/r2_dp1.json')); print('dp', d['ms_per_step'], d['roofline']['frac_of_mode_ceiling'])
PY
tail -5 gpurun_out/r2_run4.err
