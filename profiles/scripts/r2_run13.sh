cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(cd deep-q-learning_b200/csrc && timeout 200 ./gemm_tc_test 2>&1 | grep -E "kind [02] M (1024|131072)|tcgen05|ALL|FAIL")
G="python bench.py --workload dp --steps 3 --warmup 3"
$G > gpurun_out/plain_dp.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 60 --csv --log-file gpurun_out/r2_launches_dp2.csv $G > gpurun_out/ncu_dp_launches2.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_dp2.csv')) if len(r)>8]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
a=collections.OrderedDict()
for r in rows[1:]:
    name=r[ki].split('(')[0][-50:]
    v=float(r[vi].replace(',',''))
    if r[ui]=='us': v*=1e3
    elif r[ui]=='ms': v*=1e6
    x=a.setdefault(name,[0,0.0]); x[0]+=1; x[1]+=v
for k,(n,t) in sorted(a.items(), key=lambda kv:-kv[1][1])[:14]: print(f'{n:3d} {t/1e3/n:9.1f} us avg  {k}')
PY
