"""Turn an .ncu-rep into a short text summary (per-launch key metrics) for profiles/."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as fh:
        fh.write('# ncu --set full --clock-control none, report %s\n' % rep)
        for r in rows[2:]:
            fh.write('\nkernel: %s\n' % r[hdr.index('Kernel Name')])
            vals = {}
            for k in KEYS:
                if k in hdr:
                    vals[k] = (r[hdr.index(k)], units[hdr.index(k)])
                    fh.write('  %-62s %s %s\n' % (k, r[hdr.index(k)], units[hdr.index(k)]))
            try:
                t_ms = float(vals['gpu__time_duration.sum'][0]) * {'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}[vals['gpu__time_duration.sum'][1]]
                scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
                rd = float(vals['dram__bytes_read.sum'][0]) * scale[vals['dram__bytes_read.sum'][1]]
                wr = float(vals['dram__bytes_write.sum'][0]) * scale[vals['dram__bytes_write.sum'][1]]
                fh.write('  derived: dram traffic %.4f GB per launch, %.1f GB/s over the launch\n' % (
                    (rd + wr) / 1e9, (rd + wr) / (t_ms * 1e-3) / 1e9))
            except Exception as err:   # pragma: no cover
                fh.write('  derived: n/a (%s)\n' % err)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
