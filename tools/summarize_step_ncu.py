"""Per-kernel table of one profiled training step (an ncu summary for every kernel >= 1 % of the
step): launches, device time and share, DRAM bytes read + written, achieved DRAM GB/s and its fraction of the
measured HBM peak, SM / memory throughput (% of peak), tensor-pipe activity.

    ncu --section SpeedOfLight --metrics dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off \
        -o gpurun_out/step_r2 python tools/profile_step.py
    python tools/summarize_step_ncu.py gpurun_out/step_r2.ncu-rep > profiles/step_r2_kernels.txt
Times are ncu's (serialised, cold caches): compare SHARES and per-kernel bandwidth, not the absolute step time."""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6530.6
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.isfile(pk):
    peak = float(json.load(open(pk))['hbm_gbs'])
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]


def col(name):
    return hdr.index(name) if name in hdr else None


def val(r, name, scale_units=True):
    i = col(name)
    if i is None or r[i] in ('', 'n/a'):
        return 0.0
    v = float(r[i].replace(',', ''))
    if scale_units:
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3,
              'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6}.get(units[i], 1)
    return v


agg = OrderedDict()
for r in rows[2:]:
    name = r[col('Kernel Name')].split('(')[0]
    a = agg.setdefault(name, dict(n=0, us=0.0, rd=0.0, wr=0.0, sm=0.0, mem=0.0, tc=0.0))
    us = val(r, 'gpu__time_duration.sum')
    a['n'] += 1
    a['us'] += us
    a['rd'] += val(r, 'dram__bytes_read.sum')
    a['wr'] += val(r, 'dram__bytes_write.sum')
    a['sm'] += us * val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed', False)
    a['mem'] += us * val(r, 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', False)
    a['tc'] += us * val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', False)
total = sum(a['us'] for a in agg.values())
print('# %s: %d launches, %.1f us in total (ncu: serialised, cold caches); HBM peak %.1f GB/s (measured)'
      % (os.path.basename(sys.argv[1]), sum(a['n'] for a in agg.values()), total, peak))
print('%-58s %4s %9s %6s %9s %9s %8s %6s %6s %6s %6s' % ('kernel', 'n', 'us', 'share', 'dram rd MB', 'dram wr MB',
                                                        'GB/s', '/peak', 'SM %', 'mem %', 'TC %'))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    gbs = (a['rd'] + a['wr']) / (a['us'] * 1e-6) / 1e9 if a['us'] > 0 else 0.0
    mark = '' if a['us'] >= 0.01 * total else '  (< 1 %)'
    print('%-58s %4d %9.1f %5.1f%% %9.1f %9.1f %8.0f %6.3f %6.1f %6.1f %6.1f%s'
          % (name[:58], a['n'], a['us'], 100 * a['us'] / total, a['rd'] / 1e6, a['wr'] / 1e6, gbs, gbs / peak,
             a['sm'] / a['us'], a['mem'] / a['us'], a['tc'] / a['us'], mark))
