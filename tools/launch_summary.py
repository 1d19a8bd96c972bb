"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid)."""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1000 if row['Metric Unit'] == 'ns' else (v * 1000 if row['Metric Unit'] == 'ms' else v)
    k = (row['Kernel Name'].replace('bmc::<unnamed>::', '').replace('void ', '')[:58], row['Grid Size'], row['Block Size'])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print('%-60s %-16s %-14s %5s %11s %9s %6s' % ('kernel', 'grid', 'block', 'n', 'total us', 'avg us', 'share'))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-60s %-16s %-14s %5d %11.1f %9.1f %5.1f%%' % (k[0], k[1], k[2], n, t, t / n, 100 * t / tot))
print('total %.1f us over %d launches' % (tot, sum(n for n, _ in agg.values())))
