#!/bin/bash
T=${1:-r02_r}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_prb_reparam.py tests/test_shim_optim.py -m gpu -q -x) > gpurun_out/${T}_pytest_prb.log 2>&1; tail -3 gpurun_out/${T}_pytest_prb.log
(timeout 600 python scripts/time_prb.py 512 8 3; timeout 600 python scripts/time_prb.py 1024 8 3) > gpurun_out/${T}_time_prb.log 2>&1; cat gpurun_out/${T}_time_prb.log
