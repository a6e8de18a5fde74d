"""Key metrics of an .ncu-rep (first kernel instance) as a small text table.
    python tools/ncu_summary.py <file.ncu-rep> [--traffic-json profiles/gather_traffic.json]
With --traffic-json the DRAM bytes of the launch (read + written) are also written to that file, which bench.py
reads for `roofline.traffic` (so the figure always comes from a committed capture, never from a constant)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print('# %s' % sys.argv[1])
def _bytes(r, key):
    i = hdr.index(key)
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
    return float(r[i].replace(',', '')) * scale


for r in rows[2:]:
    print('kernel:', r[hdr.index('Kernel Name')][:100])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print('  %-95s %s %s' % (k, r[i], units[i]))
if '--traffic-json' in sys.argv:
    import json
    import os
    r = rows[2]
    rd, wr = _bytes(r, 'dram__bytes_read.sum'), _bytes(r, 'dram__bytes_write.sum')
    out = {'kernel': r[hdr.index('Kernel Name')], 'dram_bytes': int(rd + wr), 'dram_bytes_read': int(rd),
           'dram_bytes_write': int(wr), 'duration_us_under_ncu': float(r[hdr.index('gpu__time_duration.sum')].replace(',', '')),
           'source': '%s (ncu --set full, one cold launch after an L2 flush; outputs still in L2 at kernel end count as '
                     '0 bytes written)' % os.path.basename(sys.argv[1])}
    json.dump(out, open(sys.argv[sys.argv.index('--traffic-json') + 1], 'w'), indent=1)
