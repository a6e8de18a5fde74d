"""Gather-kernel lab: microseconds per launch of the roofline kernel variants at the BASELINE shape
(same harness as bench.py's roofline leg).  CGC_GATHER_STAGE=1|2 in the environment truncates the ELL kernel
(1 = exit at entry: launch floor; 2 = exit after the TMA tile copy)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--variants', default='ell,ell-noparity,csr')
a = ap.parse_args()
args = argparse.Namespace(batch_per_gpu=64, nodes=1800, degree=8)
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
for v in a.variants.split(','):
    variant, parity = ('ell', False) if v == 'ell-noparity' else (v, True)
    alg, us, copies = bench.gather_roofline(args, dev, variant, parity)
    print('%-14s stage=%s  %.2f us/launch  %.0f GB/s algorithmic  samples %s' % (
        v, os.environ.get('CGC_GATHER_STAGE', '0'), us, alg / us / 1e3,
        ' '.join('%.2f' % t for t in bench.gather_roofline.samples)))
