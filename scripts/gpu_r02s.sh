#!/bin/bash
T=${1:-r02_s}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_prb.csv python scripts/time_prb.py 512 8 1 > gpurun_out/${T}_ncu_list.log 2>&1
python - <<PY
import csv,re,collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/${T}_launches_prb.csv') if l.startswith('"'))]
H=rows[0]; ki=H.index('Kernel Name'); vi=H.index('Metric Value'); gi=H.index('Grid Size')
names=[(re.sub(r'\(.*','',r[ki]).replace('void ',''), r[gi], int(r[vi].replace(',',''))) for r in rows[1:]]
tot=collections.OrderedDict()
for n,g,t in names:
    tot.setdefault(n,[0,0]); tot[n][0]+=1; tot[n][1]+=t
for n,(c,t) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:8]: print("%-60s %4d %9.2f ms"%(n[:60],c,t/1e6))
print([round(t/1e6,2) for n,g,t in names if n.startswith('k_prb_pass')])
PY
