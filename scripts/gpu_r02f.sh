#!/bin/bash
T=${1:-r02_f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "configs_1_and_3" 2>&1 | tail -3
for L in libepsm_b200.so libepsm_prb3.so libepsm_prb4.so; do EPSM_B200_LIB=$PWD/epsm_mitsuba3_b200/csrc/$L timeout 600 python scripts/time_prb.py 256 8 2>&1 | tail -2; done | tee gpurun_out/${T}_time_prb.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-iteration-leg --no-full-replay-leg --workload room8m --full-replay"
timeout 900 $B > gpurun_out/${T}_bench_room8m_full.json 2> gpurun_out/${T}_bench_room8m_full.err; tail -c 900 gpurun_out/${T}_bench_room8m_full.json
NB="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-full-replay-leg --no-iteration-leg --workload room8m --full-replay"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -c 13 -o gpurun_out/${T}_trace8m_fullreplay -f $NB > gpurun_out/${T}_ncu_trace8m.log 2>&1
ncu -i gpurun_out/${T}_trace8m_fullreplay.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,dram__throughput.avg.pct_of_peak_sustained_elapsed > gpurun_out/${T}_trace8m_metrics.csv 2>&1
tail -14 gpurun_out/${T}_trace8m_metrics.csv | cut -c1-300
