#!/bin/bash
# N-GPU bench line (strong headline + weak leg), launched the way the driver does.
N=${1:-2}; T=${2:-r02_p}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-iteration-leg --no-full-replay-leg > gpurun_out/${T}_bench_room_n${N}.json 2> gpurun_out/${T}_bench_room_n${N}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_room_n${N}.json').read().strip().splitlines()[-1])
print('N=$N', 'strong ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e'] and d['e2e']['value'], 'weak', d['weak'] and (d['weak']['ms_per_step'], d['weak']['value']))
PY
tail -2 gpurun_out/${T}_bench_room_n${N}.err
