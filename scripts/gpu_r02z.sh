#!/bin/bash
# Round-2 closing measurement call: GPU tests, default bench, reference arm, other workloads, phase-2 leg, launch list, ncu --set full of the pass.
T=${1:-r02_z}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 700 python bench.py > gpurun_out/${T}_bench_room.json 2> gpurun_out/${T}_bench_room.err; tail -c 700 gpurun_out/${T}_bench_room.json; echo
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 600 gpurun_out/${T}_bench_reference.json; echo
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-iteration-leg"
for W in glass smpl cbox; do timeout 400 $B --workload $W > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err; python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_$W.json').read().strip().splitlines()[-1]); print('$W', d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'])"; done
timeout 900 $B --no-e2e --no-full-replay-leg --phase2-leg > gpurun_out/${T}_bench_phase2.json 2> gpurun_out/${T}_bench_phase2.err; python -c "
import json; d=json.loads(open('gpurun_out/${T}_bench_phase2.json').read().strip().splitlines()[-1]); print('phase2', d['phase2_iteration'])"
bash scripts/gpu_launchlist.sh $T > gpurun_out/${T}_launch_table.txt 2>&1; head -14 gpurun_out/${T}_launch_table.txt
NB="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-full-replay-leg --no-iteration-leg"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf_ -c 24 -o gpurun_out/${T}_pass_full -f $NB > gpurun_out/${T}_ncu_pass.log 2>&1
python scripts/ncu_table.py gpurun_out/${T}_pass_full.ncu-rep > gpurun_out/${T}_pass_table.txt 2>&1; head -c 600 gpurun_out/${T}_pass_table.txt
# the report itself is too large for the 64 MiB return path: keep the table and the per-kernel details of four launches
ncu -i gpurun_out/${T}_pass_full.ncu-rep --page details --csv --kernel-name regex:'k_wf_(trace_primary|finish)' > gpurun_out/${T}_pass_details.csv 2>/dev/null
rm -f gpurun_out/${T}_pass_full.ncu-rep; du -sh gpurun_out
