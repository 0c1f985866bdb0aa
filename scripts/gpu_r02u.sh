#!/bin/bash
# GPU tests of the current tree + shared-stack depth A/B (WF_STACK 16 / 12 / 8) on room (gradient pass) and the primal render
T=${1:-r02_u}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
for L in libepsm_b200.so libepsm_ws12.so libepsm_ws8.so; do
  export EPSM_B200_LIB=$PWD/epsm_mitsuba3_b200/csrc/$L
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-iteration-leg --no-full-replay-leg > gpurun_out/${T}_${L%.so}_room.json 2>/dev/null
  python - gpurun_out/${T}_${L%.so}_room.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print(sys.argv[1], 'ms/step %.2f trace_ms %.2f' % (d['ms_per_step'], r['kernel_ms_per_step']))
PY
  python scripts/time_primal.py 2>&1 | tail -1
done
unset EPSM_B200_LIB
