#!/bin/bash
# GPU tests (stop at first failure) + quick bench + primal timing
T=${1:-r02_l}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${T}_pytest_gpu.log
bash scripts/gpu_quick2.sh $T
