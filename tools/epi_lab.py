"""µs per call of the wide element-wise operators of the pool GNN's last layer (20 -> 1140 channels, 115 200 rows: the
config #3 shapes), each timed alone with CUDA events over 10 calls after a warm-up.  Operands are 525 MB each, so
every call streams from HBM.  Usage (GPU box): python tools/epi_lab.py   (CGC_EPI_RING=0 on a LAB build: old kernel)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cgc_net_b200 import _backend  # noqa: E402

dev = torch.device('cuda:0')
be = _backend.get()
b, nmax, cin, cout = 64, 1800, 20, 1140
rows = b * nmax
g = torch.Generator(device='cpu').manual_seed(0)
sa = torch.randn(rows, cin, generator=g).to(dev)
w = (torch.randn(cin, cout, generator=g) / cin ** 0.5).to(dev)
bias = (torch.randn(cout, generator=g) * 0.1).to(dev)
mask = torch.full((b,), nmax, dtype=torch.int32, device=dev), nmax
cat = torch.zeros(rows, 1180, device=dev)
y = cat[:, 40:]
gcat = torch.randn(rows, 1180, device=dev)
gx = gcat[:, 40:]
gy = torch.empty(rows, cout, device=dev)
dh = torch.empty(rows, cout, device=dev)
mean, invstd = torch.zeros(cout, device=dev), torch.ones(cout, device=dev)
rmean, rvar = torch.zeros(cout, device=dev), torch.ones(cout, device=dev)
gamma = torch.ones(cout, device=dev)
dg, db, out = (torch.zeros(cout, device=dev) for _ in range(3))


def timed(name, fn, gbytes, reps=10):
    fn()
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print('%-46s %8.1f us   %6.0f GB/s algorithmic' % (name, us, gbytes * 1e3 / us), flush=True)


from cgc_net_b200 import ops  # noqa: E402
m = ops.RowMask(mask[0], nmax)
one = rows * cout * 4 / 1e6       # MB of one wide matrix
timed('sage_layer_fwd_wide (epilogue + BN sums)', lambda: be.sage_layer_fwd_wide(sa, w, bias, m, y, 0, rmean, rvar, 0.1, 1e-5, mean, invstd), one)
timed('bn_apply', lambda: be.bn_apply(y, mean, invstd, gamma, bias, m, gy), 2 * one)
timed('bn_bwd sums only', lambda: be.bn_bwd(gx, y, mean, invstd, gamma, m, True, None, dg, db), 2 * one)
timed('bn_bwd sums + apply', lambda: be.bn_bwd(gx, y, mean, invstd, gamma, m, True, gy, dg, db), 5 * one)
timed('sage_dh (in place)', lambda: be.sage_dh(gy, sa, w, bias, y, m, gy, 0), 3 * one)
timed('sage_dh_bn (BatchNorm backward folded)', lambda: be.sage_dh_bn(gx, sa, w, bias, y, mean, invstd, gamma, dg, db, True, m, dh, 0), 3 * one)
timed('colsum', lambda: be.colsum(gy, None, out), one)
timed('bn_stats', lambda: be.bn_stats(y, rmean, rvar, 0.1, 1e-5, mean, invstd), one)
