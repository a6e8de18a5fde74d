"""Launch the tcgen05 GEMM a few times on the `lin` contraction shape (for ncu --set full and for A/B timing of
the lab switches of a LAB build).  Environment: M, N, K, PREC (1 = 3xTF32, 3 = single pass), TA / TB (operand
layouts: A stored [k, m] / B stored [n, k])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cgc_net_b200 import _backend  # noqa: E402

be = _backend.get()
m, n, k = int(os.environ.get('M', 115200)), int(os.environ.get('N', 1140)), int(os.environ.get('K', 1180))
prec = int(os.environ.get('PREC', 1))
ta, tb = bool(int(os.environ.get('TA', 0))), bool(int(os.environ.get('TB', 1)))
batch = int(os.environ.get('BATCH', 0))            # > 0: batched contraction (dense stages: BATCH=64 M=1140 N=60 K=1140 TB=0)
lead = (batch,) if batch else ()
a = torch.randn(lead + ((k, m) if ta else (m, k)), device='cuda')
b = torch.randn(lead + ((n, k) if tb else (k, n)), device='cuda')
c = torch.empty(lead + (m, n), device='cuda')
for _ in range(3):
    be.gemm(a, b, c, ta, tb, precision=prec)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    be.gemm(a, b, c, ta, tb, precision=prec)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
lab = ' '.join('%s=%s' % (k_, v) for k_, v in sorted(os.environ.items()) if k_.startswith('CGC_TC_'))
print('gemm %s%dx%dx%d prec %d ta %d tb %d %s: %.3f ms  %.1f TFLOP/s (fp32-equivalent)  A read at %.0f GB/s'
      % ('%d x ' % batch if batch else '', m, n, k, prec, ta, tb, lab, ms, 2.0 * max(batch, 1) * m * n * k / ms / 1e9,
         4.0 * max(batch, 1) * m * k / ms / 1e6))
