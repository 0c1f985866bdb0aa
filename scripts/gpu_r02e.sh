#!/bin/bash
# Round-2 measurement call: GPU tests, default bench, reference arm, 8M-triangle (HBM-resident) workload + its ncu capture, phase-2 iteration.
T=${1:-r02_e}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 700 python bench.py > gpurun_out/${T}_bench_room.json 2> gpurun_out/${T}_bench_room.err; tail -c 600 gpurun_out/${T}_bench_room.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; tail -c 1500 gpurun_out/${T}_bench_reference.json
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-iteration-leg"
timeout 600 $B --workload room8m > gpurun_out/${T}_bench_room8m.json 2> gpurun_out/${T}_bench_room8m.err; tail -c 1200 gpurun_out/${T}_bench_room8m.json
timeout 900 $B --no-full-replay-leg --phase2-leg > gpurun_out/${T}_bench_phase2.json 2> gpurun_out/${T}_bench_phase2.err; tail -c 700 gpurun_out/${T}_bench_phase2.json
NB="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-full-replay-leg --no-iteration-leg --workload room8m"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -c 12 -o gpurun_out/${T}_trace8m_full -f $NB > gpurun_out/${T}_ncu_trace8m.log 2>&1
ncu -i gpurun_out/${T}_trace8m_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio > gpurun_out/${T}_trace8m_metrics.csv 2>&1
tail -14 gpurun_out/${T}_trace8m_metrics.csv | cut -c1-400
