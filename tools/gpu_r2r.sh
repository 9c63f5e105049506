#!/bin/bash
TAG=${1:-r2r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], d['solver'], 'frac', d['roofline']['frac'])
for k,v in d.get('configs',{}).items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ('solves_per_s','kernel_ms','seconds','kernel','ok')}, v.get('roofline',{}).get('frac'))
print(d.get('cpu_baseline'))
PY
