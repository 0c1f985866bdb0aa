mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/v_default.json 2>gpurun_out/v_default.err
for c in 23 24 25; do EPSM_WF_CHUNK=$((1<<c)) $B --no-full-replay-leg > gpurun_out/v_chunk$c.json 2>gpurun_out/v_chunk$c.err; done
$B --workload glass > gpurun_out/v_default_glass.json 2>&1
$B --workload smpl > gpurun_out/v_default_smpl.json 2>&1
for f in gpurun_out/v_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']; fr=d.get('full_replay') or {}
    print(sys.argv[1], 'ms/step %.1f value %.1fM trace_ms %.1f share %.2f frac %.3f rays/path %.2f nodes/path %.1f full %.1f ms' % (d['ms_per_step'], d['value']/1e6, r['kernel_ms_per_step'], r['share_of_step'], r['frac'], r['per_path']['rays'], r['per_path']['nodes'], fr.get('ms_per_step',0)))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_room_pruned.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-full-replay-leg > gpurun_out/ncu_list2.log 2>&1
