"""Launch the tcgen05 GEMM a few times on the `lin` contraction shape (for ncu --set full)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cgc_net_b200 import _backend  # noqa: E402

be = _backend.get()
m, n, k = int(os.environ.get('M', 115200)), int(os.environ.get('N', 1140)), int(os.environ.get('K', 1180))
a = torch.randn(m, k, device='cuda')
b = torch.randn(n, k, device='cuda')
c = torch.empty(m, n, device='cuda')
for _ in range(3):
    be.gemm(a, b, c, False, True, precision=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    be.gemm(a, b, c, False, True, precision=1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('gemm %dx%dx%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent)' % (m, n, k, ms, 2.0 * m * n * k / ms / 1e9))
