#!/bin/bash
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 300 python tools/prof_step_gap.py > gpurun_out/${TAG}_gap.log 2>&1; cat gpurun_out/${TAG}_gap.log | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-large > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], d['solver']['pulse_ms'], d['solver']['relax_ms'])
for k,v in d.get('configs',{}).items():
    print(k, {kk:vv for kk,vv in v.items() if kk in ('solves_per_s','kernel_ms','seconds','kernel','ok')}, v.get('roofline',{}).get('frac'))
PY
