"""SASS opcode evidence of the release build: per kernel of libcgc_sm100.so, how many Blackwell-native
instructions it contains -- UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (cp.async.bulk.tensor),
UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus the legacy tensor path HMMA (must be 0).
Usage: python tools/sass_histogram.py > profiles/sass_r2.txt   (runs here: cuobjdump needs no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'cgc_net_b200', 'csrc', 'libcgc_sm100.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'ACQBULK',
       'HMMA', 'HGMMA', 'LDGSTS', 'FFMA2', 'FADD2', 'LDS', 'SHFL']

sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
demangle = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True,
                          text=True).stdout.splitlines()
names = iter(demangle)
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = next(names).split('(')[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m:
        op = m.group(1)
        for key in OPS:
            if op == key or op.startswith(key + '.') or op.startswith(key):
                counts[cur][key] += 1
                break
arch = re.findall(r'arch = (\S+)', sass)
print('# %s  (%s; %d kernels)' % (os.path.relpath(LIB, ROOT), ', '.join(sorted(set(arch))), len(counts)))
print('# kernels with tensor-core / TMA / bulk-copy / mbarrier instructions; HMMA (legacy mma.sync) must not appear')
print('%-64s %s' % ('kernel', ' '.join('%7s' % o for o in OPS)))
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    if any(c[o] for o in OPS[:13]):
        print('%-64s %s' % (k[:64], ' '.join('%7d' % c[o] for o in OPS)))
print('%-64s %s' % ('TOTAL (all kernels)', ' '.join('%7d' % tot[o] for o in OPS)))
assert tot['HMMA'] == 0 and tot['HGMMA'] == 0, 'legacy tensor-core instructions found'
