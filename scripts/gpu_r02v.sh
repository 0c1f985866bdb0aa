#!/bin/bash
# prb_reparam with contiguous auxiliary-ray queue entries (tests + timing + launch list); finish-kernel occupancy A/B
T=${1:-r02_v}
mkdir -p gpurun_out
bash scripts/gpu_r02r.sh $T
bash scripts/gpu_r02s.sh $T
for L in libepsm_b200.so libepsm_fs3.so; do
  EPSM_B200_LIB=$PWD/epsm_mitsuba3_b200/csrc/$L python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-iteration-leg --no-full-replay-leg > gpurun_out/${T}_${L%.so}_room.json 2>/dev/null
  python - gpurun_out/${T}_${L%.so}_room.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print(sys.argv[1], 'ms/step %.2f trace_ms %.2f' % (d['ms_per_step'], r['kernel_ms_per_step']))
PY
done
