#!/bin/bash
# ncu --set full of one gradient pass of the default workload → table + details of the traversal / finish launches (the report is deleted: too large to return)
T=${1:-r02_zz}
mkdir -p gpurun_out
NB="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-full-replay-leg --no-iteration-leg"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf_ -c 24 -o gpurun_out/${T}_pass_full -f $NB > gpurun_out/${T}_ncu_pass.log 2>&1
python scripts/ncu_table.py gpurun_out/${T}_pass_full.ncu-rep > gpurun_out/${T}_pass_table.txt 2>&1; head -c 400 gpurun_out/${T}_pass_table.txt
rm -f gpurun_out/${T}_pass_full.ncu-rep
