#!/bin/bash
# A/B of library variants on the room pass: scripts/gpu_ab2.sh <tag> <lib.so> …
T=$1; shift
mkdir -p gpurun_out
for L in "$@"; do
  EPSM_B200_LIB=$PWD/epsm_mitsuba3_b200/csrc/$L python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-iteration-leg --no-full-replay-leg > gpurun_out/${T}_${L%.so}_room.json 2>/dev/null
  python - gpurun_out/${T}_${L%.so}_room.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print(sys.argv[1], 'ms/step %.2f trace_ms %.2f other %.2f' % (d['ms_per_step'], r['kernel_ms_per_step'], d['ms_per_step']-r['kernel_ms_per_step']))
PY
done
