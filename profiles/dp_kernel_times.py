import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10 and r[0].isdigit()]
# find the last lb_adam_kernel and take the launches between the previous adam and it
names=[r[4] for r in rows]
adam=[i for i,n in enumerate(names) if 'lb_adam_kernel' in n]
seg=rows[adam[-2]+1:adam[-1]+1]
d=collections.OrderedDict()
for r in seg:
    n=r[4].split("(")[0].replace("void ","").replace("dqn::<unnamed>::","").replace("dqn::","")[:40]+" "+r[8]
    d[n]=d.get(n,0)+float(r[-1])
tot=sum(d.values())
for k,v in sorted(d.items(), key=lambda kv:-kv[1])[:12]: print("%-60s %8.1f us %5.1f%%"%(k,v/1e3,100*v/tot))
print("one step under ncu: %.3f ms" % (tot/1e6))
