#!/bin/bash
# Round-2 step J: compressed 8-wide BVH.  GPU tests, default bench, primal timing, full-replay 8M-triangle leg.
T=${1:-r02_j}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 700 python bench.py > gpurun_out/${T}_bench_room.json 2> gpurun_out/${T}_bench_room.err; tail -c 2500 gpurun_out/${T}_bench_room.json; tail -3 gpurun_out/${T}_bench_room.err
timeout 300 python scripts/time_primal.py > gpurun_out/${T}_time_primal.log 2>&1; tail -2 gpurun_out/${T}_time_primal.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-iteration-leg"
timeout 600 $B --workload room8m --full-replay > gpurun_out/${T}_bench_room8m_full.json 2> gpurun_out/${T}_bench_room8m_full.err; tail -c 1500 gpurun_out/${T}_bench_room8m_full.json
