"""Write profiles/r02_traffic.json: DRAM bytes per launch of the dominant kernels from `.ncu-rep`
captures (`ncu --set full`).  Usage: python tools/ncu_traffic.py <cluster.ncu-rep> [<tma.ncu-rep>]"""
import csv
import json
import subprocess
import sys

SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}


def rows_of(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rd = float(d['dram__bytes_read.sum']) * SCALE[u['dram__bytes_read.sum']]
        wr = float(d['dram__bytes_write.sum']) * SCALE[u['dram__bytes_write.sum']]
        out.append((d['Kernel Name'], rd + wr, float(d['gpu__time_duration.sum']), u['gpu__time_duration.sum']))
    return out


def main():
    out = {}
    cl = rows_of(sys.argv[1])
    name, traffic, t, tu = cl[-1]
    out['k_integrate_cl_relax_dram_bytes'] = traffic
    out['source'] = ('profiles/%s: ncu --set full --clock-control none of the relaxation launch of the C2 batch '
                     '(32 systems, 1000 output times): dram read+write = %.1f MB (the trajectory output is '
                     '32 x 1000 x N x 8 B = %s), %s %s under ncu' % (
                         sys.argv[1].split('/')[-1].replace('.ncu-rep', '.txt'), traffic / 1e6, 'see file', t, tu))
    if len(sys.argv) > 2:
        for name, traffic, t, tu in rows_of(sys.argv[2]):
            key = 'k_op_tma_jv_dram_bytes' if 'true' in name or 'Lb1' in name else 'k_op_tma_rhs_dram_bytes'
            out[key] = traffic
        out['k_op_tma_source'] = 'profiles/%s (ncu --set full of k_op_tma: dram read+write per launch)' % \
            sys.argv[2].split('/')[-1].replace('.ncu-rep', '.txt')
    with open('profiles/r02_traffic.json', 'w') as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
