"""tools/ncu_summary.py <launches.csv> -- aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list
by kernel: launches, total time, share.  Output is markdown (committed under profiles/)."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6, 's': 1e9, 'second': 1e9}.get(unit, 1)
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        rows.append((name, ns))
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    total = sum(v[1] for v in agg.values())
    print('| kernel | launches | total ms | share | avg us |')
    print('|---|---:|---:|---:|---:|')
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.3f | %.1f%% | %.1f |' % (n, c, ns / 1e6, 100 * ns / total, ns / c / 1e3))
    print('| **total** | %d | %.3f | 100%% | |' % (len(rows), total / 1e6))


if __name__ == '__main__':
    main(sys.argv[1])
