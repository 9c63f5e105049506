#!/bin/bash
TAG=${1:-r2n2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/${TAG}_bench.err | cut -c1-300
python - <<PY
import json
for line in open('gpurun_out/${TAG}_bench.json'):
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], 'per_rank', d['per_rank_ms'])
        print('strong', d.get('strong'))
PY
