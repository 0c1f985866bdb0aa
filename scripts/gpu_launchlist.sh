#!/bin/bash
# ncu launch list (gpu__time_duration) of the default workload's gradient pass
T=${1:-r02_m}
mkdir -p gpurun_out
NB="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-full-replay-leg --no-iteration-leg"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_room.csv $NB > gpurun_out/${T}_ncu_list.log 2>&1
python scripts/launch_table.py gpurun_out/${T}_launches_room.csv
