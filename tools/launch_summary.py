"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import defaultdict


def main(path, out, note=''):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[12] == 'gpu__time_duration.sum']
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        name = r[4].split('(')[0]
        scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}[r[13]]
        tot[name] += float(r[14].replace(',', '')) * scale
        cnt[name] += 1
    total = sum(tot.values())
    with open(out, 'w') as fh:
        fh.write('# %s\n# (cold-cache, serialised launches: compare SHARES, not absolutes); %d launches\n' % (note, len(rows)))
        fh.write('%-64s %6s %12s %7s\n' % ('kernel', 'count', 'total ms', 'share'))
        for k in sorted(tot, key=tot.get, reverse=True):
            fh.write('%-64s %6d %12.3f %6.1f%%\n' % (k[:64], cnt[k], tot[k], 100 * tot[k] / total))


if __name__ == '__main__':
    main(*sys.argv[1:])
