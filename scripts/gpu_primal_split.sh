#!/bin/bash
# per-kernel split of the primal render (ncu launch list) for a library variant: scripts/gpu_primal_split.sh <tag> <lib.so>
T=$1; L=$2
mkdir -p gpurun_out
cat > /tmp/primal_once.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, epsm_mitsuba3_b200 as ep
from epsm_mitsuba3_b200 import scenes as S
sd = S.indoor_room(1000000, 1024, 32, 1024); scene = ep.Scene(sd, device=0); integ = ep.load_dict({"type": "manifold", "max_depth": 8})
for it in range(2):
    integ.render(scene, sensor=1, seed=it, spp=32); scene.synchronize()
PY
EPSM_B200_LIB=$PWD/epsm_mitsuba3_b200/csrc/$L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_primal.csv python /tmp/primal_once.py > /dev/null 2>&1
python - gpurun_out/${T}_launches_primal.csv <<'PY'
import csv,re,sys,collections
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
H=rows[0]; ki=H.index('Kernel Name'); vi=H.index('Metric Value')
names=[(re.sub(r'\(.*','',r[ki]).replace('void ',''), int(r[vi].replace(',',''))) for r in rows[1:]]
idx=[i for i,n in enumerate(names) if n[0].startswith('k_wf_trace_primary')]
a=idx[-1]; tot=collections.OrderedDict()
for n,t in names[a:]:
    tot.setdefault(n,[0,0]); tot[n][0]+=1; tot[n][1]+=t
for n,(c,t) in tot.items(): print(sys.argv[1][-40:], '%-28s %3d %8.2f ms' % (n[:28], c, t/1e6))
PY
