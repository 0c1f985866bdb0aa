#!/bin/bash
# quick A/B call: default-bench numbers without the side legs + primal timing
T=${1:-r02_q}
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-iteration-leg --no-full-replay-leg"
timeout 600 $B > gpurun_out/${T}_bench_room.json 2> gpurun_out/${T}_bench_room.err; python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_room.json').read().strip().splitlines()[-1])
r=d['roofline']; print('${T}', 'ms_per_step', d['ms_per_step'], 'trace ms', r['kernel_ms_per_step'], 'per_path', r['per_path'], 'nodes', d['scene']['bvh_nodes'])
PY
timeout 300 python scripts/time_primal.py 2>&1 | tail -1
