#!/bin/bash
# BASELINE configs[4]: resolution / spp sweep of the gradient pass on the room scene (one GPU).  usage: scripts/sweep.sh <tag>
T=${1:-sweep}
mkdir -p gpurun_out
: > gpurun_out/${T}_sweep.jsonl
for res in 512 1024 2048; do for spp in 16 64 256; do
  timeout 600 python bench.py --res $res --spp $spp --steps 3 --warmup 3 --no-cpu-baseline --no-full-replay-leg --no-iteration-leg 2>/dev/null | tail -1 >> gpurun_out/${T}_sweep.jsonl
done; done
python - "$T" <<'PY'
import json,sys
for l in open(f'gpurun_out/{sys.argv[1]}_sweep.jsonl'):
    try:
        d=json.loads(l); c=d['config']; print(c['resolution'][0], c['spp'], round(d['ms_per_step'],2), round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), round(d['roofline']['share_of_step'],3))
    except Exception as e: print('ERR', e)
PY
