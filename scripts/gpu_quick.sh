#!/bin/bash
# One short gpurun call: GPU parity tests + the default bench line (+ optional extra command).
# usage: scripts/gpu_quick.sh <tag> [extra command ...]      (outputs under gpurun_out/<tag>_*)
T=${1:-q}; shift
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q ${PYTEST_EXTRA}) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${T}_pytest_gpu.log
timeout 700 python bench.py > gpurun_out/${T}_bench_room.json 2> gpurun_out/${T}_bench_room.err; tail -c 1500 gpurun_out/${T}_bench_room.json; tail -3 gpurun_out/${T}_bench_room.err
if [ $# -gt 0 ]; then bash -c "$*"; fi
