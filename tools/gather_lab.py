"""Gather-kernel lab: microseconds per launch of the roofline kernel family at the BASELINE shape (same harness as
bench.py's roofline leg).  With a LAB build (make LAB=1) CGC_GATHER_STAGE=1|2 truncates the ELL kernel
(1 = exit at entry: launch floor; 2 = exit after the TMA tile copy)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

VARIANTS = {
    'ell': dict(),
    'ell-nopdl': dict(pdl=False),
    'ell-1persm': dict(two_per_sm=False),
    'ell-nomc': dict(multicast=False),
    'ell-2persm-nopdl': dict(two_per_sm=True, pdl=False),
    'ell-vardeg': dict(row_is_centre=False),
    'ell-weighted': dict(norm_adj=True),
    'layer': dict(kind='layer'),
    'layer-nopdl': dict(kind='layer', pdl=False),
    'layer-weighted': dict(kind='layer', norm_adj=True),
    'csr': dict(kind='csr'),
}
ap = argparse.ArgumentParser()
ap.add_argument('--variants', default='ell,ell-nopdl,ell-1persm,ell-2persm-nopdl,ell-vardeg,ell-weighted,layer,layer-nopdl,csr')
a = ap.parse_args()
args = argparse.Namespace(batch_per_gpu=64, nodes=1800, degree=8)
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
lab = bool(os.environ.get('CGC_GATHER_STAGE'))
for v in a.variants.split(','):
    alg, us, copies = bench.gather_roofline(args, dev, check=not lab, **VARIANTS[v])
    print('%-16s stage=%s  %.2f us/launch  %.0f GB/s algorithmic (%.3f of 6530.6)  samples %s' % (
        v, os.environ.get('CGC_GATHER_STAGE', '0'), us, alg / us / 1e3, alg / us / 1e3 / 6530.6,
        ' '.join('%.2f' % t for t in bench.gather_roofline.samples)))
