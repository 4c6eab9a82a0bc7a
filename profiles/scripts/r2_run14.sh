cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest14.log
tail -5 gpurun_out/r2_pytest14.log
python bench.py --workload dp --steps 30 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dp', d['ms_per_step'], d['roofline']['frac_of_mode_ceiling'], d['clocks'])"
G="python bench.py --workload dp --steps 3 --warmup 3"
$G > gpurun_out/plain_dp.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 60 --csv --log-file gpurun_out/r2_launches_dp3.csv $G > gpurun_out/ncu_dp_launches3.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_dp3.csv')) if len(r)>8]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
a=collections.OrderedDict()
for r in rows[1:]:
    name=r[ki].split('(')[0][-50:]
    v=float(r[vi].replace(',',''))
    if r[ui]=='us': v*=1e3
    elif r[ui]=='ms': v*=1e6
    x=a.setdefault(name,[0,0.0]); x[0]+=1; x[1]+=v
for k,(n,t) in sorted(a.items(), key=lambda kv:-kv[1][1])[:9]: print(f'{n:3d} {t/1e3/n:9.1f} us avg  {k}')
PY
