#!/bin/bash
T=${1:-r02_g}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_prb_reparam.py tests/test_shim_optim.py tests/test_matcher.py -m gpu -q) > gpurun_out/${T}_pytest_prb.log 2>&1; tail -4 gpurun_out/${T}_pytest_prb.log
(timeout 300 python scripts/time_prb.py 256 8; EPSM_PRB_MEGAKERNEL=1 timeout 300 python scripts/time_prb.py 256 8; timeout 600 python scripts/time_prb.py 1024 8) 2>&1 | grep "prb backward" | tee gpurun_out/${T}_time_prb.log
(EPSM_SINK_REORDER=1 timeout 600 python scripts/bench_matcher.py 128 256; EPSM_SINK_REORDER=0 timeout 600 python scripts/bench_matcher.py 128 256) 2>&1 | tail -2 | tee gpurun_out/${T}_matcher.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-iteration-leg --no-full-replay-leg --phase2-leg"
timeout 900 $B > gpurun_out/${T}_bench_phase2.json 2> gpurun_out/${T}_bench_phase2.err; tail -c 700 gpurun_out/${T}_bench_phase2.json
