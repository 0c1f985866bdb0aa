#!/bin/bash
# A/B of library build variants on one box: scripts/gpu_ab.sh <tag> <lib.so> [<lib.so> …]  (libs under epsm_mitsuba3_b200/csrc)
T=$1; shift
mkdir -p gpurun_out
for L in "$@"; do
  for W in room glass smpl; do
    EPSM_B200_LIB=$PWD/epsm_mitsuba3_b200/csrc/$L python bench.py --steps 5 --warmup 3 --workload $W --no-cpu-baseline --no-e2e --no-iteration-leg --no-full-replay-leg \
      > gpurun_out/${T}_${L%.so}_$W.json 2> gpurun_out/${T}_${L%.so}_$W.err
    python - gpurun_out/${T}_${L%.so}_$W.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print(sys.argv[1], 'ms/step %.2f value %.1fM trace_ms %.2f other %.2f' % (d['ms_per_step'], d['value']/1e6, r['kernel_ms_per_step'], d['ms_per_step']-r['kernel_ms_per_step']))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
  done
done
