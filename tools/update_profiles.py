"""Copy the evidence of `bash tools/gpu_final.sh r02 a` from gpurun_out/ into profiles/ and print the numbers the
documents quote.  DEVELOPMENT TOOL."""
import csv
import json
import shutil
import subprocess
subprocess.run(['python', 'tools/ncu_summary.py', 'gpurun_out/r02_cl.ncu-rep', 'profiles/r02_ncu_integrate_cl.txt'], check=True)
subprocess.run(['python', 'tools/launch_summary.py', 'gpurun_out/r02_launches.csv', 'profiles/r02_ncu_launches_bench.txt'], check=True)
shutil.copy('gpurun_out/r02_launches.csv', 'profiles/r02_ncu_launches_bench.csv')
shutil.copy('gpurun_out/r02_bench.json', 'profiles/r02_bench_n1.json')
shutil.copy('gpurun_out/r02_bench_ref.json', 'profiles/r02_bench_ref_n1.json')
raw = subprocess.run(['ncu', '-i', 'gpurun_out/r02_cl.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
d, u = dict(zip(hdr, rows[-1])), dict(zip(hdr, units))
SC = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
tr = sum(float(d[k]) * SC[u[k]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
t = json.load(open('profiles/r02_traffic.json'))
t['k_integrate_cl_relax_dram_bytes'] = tr
json.dump(t, open('profiles/r02_traffic.json', 'w'), indent=1)
p = 'profiles/r02_ncu_launches_bench.txt'
s = open(p).read().split('\n')
s[0] = ('# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu '
        '--no-configs (round 2 final build)')
open(p, 'w').write('\n'.join(s))
b = json.load(open('gpurun_out/r02_bench.json'))
print('value', b['value'], 'e2e', b['e2e']['value'], 'clocks', b['clocks'])
for k, c in b['configs'].items():
    print(k, {x: c[x] for x in ('solves_per_s', 'kernel_ms', 'seconds', 'jv_passes') if x in c}, c.get('roofline', {}).get('frac'))
print('ref', json.load(open('gpurun_out/r02_bench_ref.json'))['value'])
