#!/bin/bash
# One gpurun call: GPU parity tests, bench lines, ncu launch list, ncu --set full captures of the two largest kernels.
# usage: scripts/gpu_round.sh <tag>      (outputs under gpurun_out/<tag>_*)
T=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${T}_bench_room.json 2> gpurun_out/${T}_bench_room.err
B="timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-full-replay-leg --no-iteration-leg"
for c in 23 24; do EPSM_WF_CHUNK=$((1<<c)) $B > gpurun_out/${T}_chunk$c.json 2> gpurun_out/${T}_chunk$c.err; done
$B --workload glass > gpurun_out/${T}_bench_glass.json 2>&1
$B --workload smpl > gpurun_out/${T}_bench_smpl.json 2>&1
for f in gpurun_out/${T}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']; fr=d.get('full_replay') or {}; e=d.get('e2e') or {}; c=d.get('cpu_baseline') or {}
    print(sys.argv[1], 'ms/step %.1f value %.1fM trace_ms %.1f share %.2f frac %.3f rays/path %.2f nodes/path %.1f full %.1f ms e2e %.1fM cpu %s' % (d['ms_per_step'], d['value']/1e6, r['kernel_ms_per_step'], r['share_of_step'], r['frac'], r['per_path']['rays'], r['per_path']['nodes'], fr.get('ms_per_step',0), (e.get('value') or 0)/1e6, c.get('value')))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
NB="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-full-replay-leg --no-iteration-leg"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_room.csv $NB > gpurun_out/${T}_ncu_list.log 2>&1
# full captures: the traversal launches of one chunk (warm-up pass) and two finish launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -c 12 -o gpurun_out/${T}_trace_full -f $NB > gpurun_out/${T}_ncu_trace.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_finish -c 2 -o gpurun_out/${T}_finish_full -f $NB > gpurun_out/${T}_ncu_finish.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wf_stage_b -c 2 -o gpurun_out/${T}_stageb_full -f $NB > gpurun_out/${T}_ncu_stageb.log 2>&1
ls -la gpurun_out | tail -20
