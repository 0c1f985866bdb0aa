#!/bin/bash
# BASELINE configs[4] on N GPUs: resolution / spp sweep of the gradient pass on the room scene, strong scaling (the job's samples
# are split over the ranks) with the weak-scaling leg beside it.  usage: scripts/sweep_multi.sh <N> <tag>
N=${1:-8}; T=${2:-sweep}
mkdir -p gpurun_out
: > gpurun_out/${T}_sweep_n${N}.jsonl
for cfg in "512 64" "1024 32" "1024 128" "2048 32"; do set -- $cfg
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --res $1 --spp $2 --steps 3 --warmup 3 --no-cpu-baseline --no-full-replay-leg --no-iteration-leg 2>/dev/null | tail -1 >> gpurun_out/${T}_sweep_n${N}.jsonl
done
python - "$T" "$N" <<'PY'
import json,sys
for l in open(f'gpurun_out/{sys.argv[1]}_sweep_n{sys.argv[2]}.jsonl'):
    try:
        d=json.loads(l); c=d['config']; w=d.get('weak') or {}
        print('N', d['n_gpus'], 'res', c['resolution'][0], 'spp', c['spp'], 'strong ms', round(d['ms_per_step'],2), 'M/s', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'weak M/s', round(w.get('value',0)/1e6,1))
    except Exception as e: print('ERR', e, l[:200])
PY
