#!/bin/bash
# Last lines of the round on the final tree: GPU tests, default bench, phase-2 leg, smpl iteration (configs[3], matcher at 256^2)
T=${1:-r02_zz}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 700 python bench.py > gpurun_out/${T}_bench_room.json 2> gpurun_out/${T}_bench_room.err; python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_room.json').read().strip().splitlines()[-1])
print('room', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'iter', d['iteration']['refit'], 'full', d['full_replay']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])
PY
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-full-replay-leg"
timeout 900 $B --no-iteration-leg --phase2-leg > gpurun_out/${T}_bench_phase2.json 2> gpurun_out/${T}_bench_phase2.err; python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_phase2.json').read().strip().splitlines()[-1]); print('phase2', d['phase2_iteration'])"
timeout 600 $B --workload smpl > gpurun_out/${T}_bench_smpl.json 2> gpurun_out/${T}_bench_smpl.err; python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_smpl.json').read().strip().splitlines()[-1]); print('smpl', d['ms_per_step'], d['value'], d['iteration'] and d['iteration']['refit'], d['iteration'] and d['iteration'].get('matcher'))"
