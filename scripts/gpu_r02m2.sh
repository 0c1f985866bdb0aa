#!/bin/bash
# matcher with the recentred dot-product inner loop: parity tests (64² CPU float64 oracle, 128² CUDA oracle) + timings
T=${1:-r02_aa}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_matcher.py -m gpu -q) > gpurun_out/${T}_pytest_matcher.log 2>&1; tail -4 gpurun_out/${T}_pytest_matcher.log
timeout 600 python scripts/bench_matcher.py > gpurun_out/${T}_matcher.log 2>&1; cat gpurun_out/${T}_matcher.log | tail -12
