"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for row in r:
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    unit = row.get('Metric Unit', 'ns')
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6}.get(unit, 1e-3)
    name = row['Kernel Name'].split('(')[0]
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print('total %.1f us over %d launches' % (total, sum(cnt.values())))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print('%9.1f us %5.1f%% %4d x %8.1f us  %s' % (v, 100 * v / total, cnt[k], v / cnt[k], k[:110]))
