"""ctypes binding of ``libcgc_sm100.so`` (the C-ABI declared in ``include/cgc_sm100.h``).

``get()`` returns the one backend the product path uses: ``CudaBackend``.  There is NO
CPU / eager fallback -- if the shared library has not been built, or a tensor is not on a
CUDA device, the call raises.  ``use_backend_for_testing`` exists only so ``tests/`` can
substitute the per-operator CPU formulas of ``oracle/op_reference.py`` to check the autograd
wiring of ``ops.py`` on a host without a GPU; nothing in the package calls it.

Each method is a 1:1 wrapper of the ``extern "C"`` function of the same name
(``cgc_<method>``): it extracts raw device pointers / leading dimensions from the torch
tensors, passes the current CUDA stream, and raises ``CgcError`` on a non-zero return code.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libcgc_sm100.so')

_c_i32, _c_i64, _c_f32, _c_vp, _c_int = (ctypes.c_int32, ctypes.c_int64, ctypes.c_float,
                                         ctypes.c_void_p, ctypes.c_int)
_c_f64 = ctypes.c_double

# name -> argtypes (every function returns int unless listed in _RESTYPES)
_SIGNATURES = {
    'cgc_version': [],
    'cgc_last_error': [],
    'cgc_workspace_bytes': [],
    'cgc_launch_count': [],
    'cgc_count_nodes': [_c_vp, _c_i64, _c_i32, _c_vp, _c_vp],
    'cgc_pad_nodes': [_c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_i32, _c_vp, _c_vp],
    'cgc_coo_to_csr_workspace': [_c_i64, _c_i64],
    'cgc_coo_to_csr': [_c_vp, _c_i64, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_int, _c_vp, _c_vp,
                       _c_vp, _c_i64, _c_vp],
    'cgc_row_inv_degree': [_c_vp, _c_i64, _c_vp, _c_vp],
    'cgc_edge_renorm': [_c_vp, _c_vp, _c_i64, _c_f64, _c_vp, _c_vp, _c_vp],
    'cgc_csr_transpose': [_c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp],
    'cgc_radius_graph': [_c_vp, _c_vp, _c_i32, _c_i32, _c_f64, _c_i32, _c_int, _c_vp, _c_vp, _c_vp],
    'cgc_knn_to_coo': [_c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i64, _c_vp, _c_vp],
    'cgc_farthest_point_sample': [_c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_i32, _c_vp, _c_vp],
    'cgc_spmm': [_c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_i32, _c_vp, _c_vp, _c_vp, _c_i64,
                 _c_int, _c_vp],
    'cgc_sage_gather_mean_fwd': [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_i32, _c_i32, _c_vp,
                                 _c_vp],
    'cgc_ell16_rows': [_c_i32, _c_i32],
    'cgc_csr_to_ell16': [_c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_i32, _c_i32, _c_int, _c_vp, _c_vp, _c_vp, _c_vp,
                         _c_vp],
    'cgc_sage_gather_ell_fwd': [_c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_i32,
                                _c_i32, _c_vp, _c_vp],
    'cgc_sage_gather_ell': [_c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32,
                            _c_i32, _c_vp, _c_i64, _c_int, _c_vp],
    'cgc_sage_layer_ell_supported': [_c_i32, _c_i32, _c_i32],
    'cgc_sage_layer_ell_fwd': [_c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32,
                               _c_i32, _c_vp, _c_vp, _c_vp, _c_i32, _c_int, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp,
                               _c_f32, _c_f32, _c_vp, _c_vp, _c_vp, _c_i64, _c_int, _c_vp],
    'cgc_gather_wide_ell': [_c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32,
                            _c_i32, _c_vp, _c_i64, _c_int, _c_vp],
    'cgc_sage_epilogue_fwd': [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i32, _c_i64, _c_i32, _c_i32, _c_int, _c_vp, _c_i64, _c_vp],
    'cgc_sage_layer_fwd_narrow': [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i32, _c_i64, _c_i32, _c_i32, _c_int, _c_vp,
                                  _c_i64, _c_vp, _c_vp, _c_f32, _c_f32, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp],
    'cgc_sage_layer_fwd_wide': [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i32, _c_i64, _c_i32, _c_i32, _c_int, _c_vp,
                                _c_i64, _c_vp, _c_vp, _c_f32, _c_f32, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp],
    'cgc_sage_dh': [_c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_i32, _c_i64, _c_i32, _c_i32, _c_int, _c_vp, _c_i64, _c_vp],
    'cgc_sage_dh_bn': [_c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int,
                       _c_vp, _c_i32, _c_i64, _c_i32, _c_i32, _c_int, _c_vp, _c_i64, _c_vp],
    'cgc_sage_layer_bwd_narrow': [_c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp, _c_i32, _c_i64, _c_i32, _c_i32, _c_int, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp],
    'cgc_act_fwd': [_c_vp, _c_i64, _c_i64, _c_i32, _c_int, _c_vp, _c_i64, _c_vp],
    'cgc_act_bwd': [_c_vp, _c_i64, _c_vp, _c_i64, _c_i64, _c_i32, _c_int, _c_vp, _c_i64, _c_vp],
    'cgc_jk_supported': [_c_i32, _c_i32],
    'cgc_jk_tape_floats': [_c_i64, _c_i32, _c_i32],
    'cgc_jk_lstm_fwd': [_c_vp, _c_i64, _c_i64, _c_i32, _c_i32] + [_c_vp] * 10 + [_c_vp, _c_i64, _c_vp, _c_vp],
    'cgc_jk_lstm_bwd': [_c_vp, _c_i64, _c_vp, _c_i64, _c_i64, _c_i32, _c_i32] + [_c_vp] * 6 +
                       [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp],
    'cgc_scale_rows_bwd': [_c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_vp, _c_i64, _c_vp, _c_vp],
    'cgc_bn_stats': [_c_vp, _c_i64, _c_i64, _c_i32, _c_vp, _c_vp, _c_f32, _c_f32, _c_vp, _c_vp,
                     _c_vp, _c_i64, _c_vp],
    'cgc_bn_eval_stats': [_c_vp, _c_vp, _c_i32, _c_f32, _c_vp, _c_vp, _c_vp],
    'cgc_bn_apply': [_c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_i64, _c_i32,
                     _c_vp, _c_i64, _c_vp],
    'cgc_bn_bwd': [_c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_i32, _c_i64,
                   _c_i32, _c_int, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp],
    'cgc_softmax_fwd': [_c_vp, _c_i64, _c_i32, _c_vp, _c_i32, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp],
    'cgc_softmax_bwd': [_c_vp, _c_vp, _c_i64, _c_i64, _c_i32, _c_vp, _c_i64, _c_vp],
    'cgc_colmax_fwd': [_c_vp, _c_i32, _c_i32, _c_i32, _c_vp, _c_vp, _c_vp],
    'cgc_colmax_bwd': [_c_vp, _c_vp, _c_i32, _c_i32, _c_i32, _c_vp, _c_vp],
    'cgc_row_inv_clamped_sum': [_c_vp, _c_i64, _c_i32, _c_vp, _c_vp, _c_vp],
    'cgc_add_rs_grad': [_c_vp, _c_vp, _c_vp, _c_i64, _c_i32, _c_vp],
    'cgc_dense_renorm_fwd': [_c_vp, _c_i32, _c_i32, _c_f64, _c_vp, _c_vp, _c_vp, _c_vp],
    'cgc_dense_renorm_bwd': [_c_vp, _c_vp, _c_vp, _c_i32, _c_i32, _c_f64, _c_vp, _c_vp, _c_vp],
    'cgc_gemm': [_c_int, _c_int, _c_i32, _c_i32, _c_i32, _c_f32, _c_vp, _c_i64, _c_i64, _c_vp,
                 _c_i64, _c_i64, _c_f32, _c_vp, _c_i64, _c_i64, _c_i32, _c_vp, _c_vp, _c_vp,
                 _c_i32, _c_int, _c_vp, _c_i64, _c_vp],
    'cgc_colsum': [_c_vp, _c_i64, _c_i64, _c_i32, _c_vp, _c_i32, _c_vp, _c_vp, _c_i64, _c_vp],
}
_RESTYPES = {'cgc_last_error': ctypes.c_char_p, 'cgc_workspace_bytes': _c_i64, 'cgc_launch_count': _c_i64,
             'cgc_coo_to_csr_workspace': _c_i64, 'cgc_ell16_rows': _c_i64, 'cgc_jk_tape_floats': _c_i64}
EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


GATHER_ACCUMULATE, GATHER_TABLES_READY, GATHER_NO_PDL, GATHER_TWO_PER_SM, GATHER_ONE_PER_SM = 1, 2, 4, 8, 16   # CGC_GATHER_*
GATHER_NO_MULTICAST = 32


class CgcError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    """dlopen the library and attach argtypes; raises if it is missing or a symbol is absent."""
    if not os.path.isfile(path):
        raise CgcError('%s not found: build it with `python -c "import __graft_entry__ as g; '
                       'g.build()"` (or `make -C cgc_net_b200/csrc`). There is no CPU fallback.' % path)
    lib = ctypes.CDLL(path)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError -> missing export
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _c_int)
    return lib


def _ptr(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise CgcError('libcgc_sm100 needs CUDA tensors (got a %s tensor); no CPU fallback' % t.device)
    if t.device.index != torch.cuda.current_device():
        raise CgcError('tensor on %s but the current CUDA device is cuda:%d: the library enqueues on the current '
                       'device (use `with torch.cuda.device(t.device):`)' % (t.device, torch.cuda.current_device()))
    if dtype is not None and t.dtype != dtype:
        raise CgcError('expected %s, got %s' % (dtype, t.dtype))
    return t.data_ptr()


def _mat(t):
    """(pointer, leading dimension) of a 2-D fp32 tensor whose last dim is contiguous."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise CgcError('expected a row-major 2-D tensor, got shape %s strides %s'
                       % (tuple(t.shape), t.stride()))
    return _ptr(t, torch.float32), t.stride(0)


def _mask(mask):
    if mask is None:
        return None, 0
    return _ptr(mask.num_nodes, torch.int32), mask.nmax


class CudaBackend(object):
    name = 'libcgc_sm100'

    def __init__(self, path=LIB_PATH):
        self.lib = load_library(path)
        self._work = {}
        self.work_bytes = int(self.lib.cgc_workspace_bytes())
        self.launches = 0           # C-ABI calls issued (bench.py reports it as gpu_launches)
        # GEMM mode of calls that do not name one: -1 = library's choice (tcgen05 3xTF32 where the shape allows),
        # 0 = exact-fp32 FMA everywhere, 1 = tcgen05 forced.  A Python attribute only (tests flip it); the
        # library itself reads no environment variable.
        self.gemm_precision = -1

    def kernel_launches(self):
        """CUDA kernels launched by the library so far (process-wide)."""
        return int(self.lib.cgc_launch_count())

    # -- helpers
    def _stream(self):
        """Stream the kernels are enqueued on.  The C entry points use the calling thread's CURRENT device
        (include/cgc_sm100.h, "Threading"), so every tensor of a call must live on it: ``_ptr`` checks that and
        raises instead of launching on foreign pointers (e.g. a model on cuda:1 while cuda:0 is current --
        wrap the call in ``torch.cuda.device(...)`` as ``compat.DataParallel`` and ``bench.py`` do)."""
        return torch.cuda.current_stream().cuda_stream

    def _workspace(self, device):
        """One workspace per (device, stream): split-K partials, column-reduction partials and the completion
        counters of two streams must not share scratch memory (kernels on one stream are ordered, across streams
        they are not)."""
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        w = self._work.get(key)
        if w is None:
            w = torch.empty(self.work_bytes, dtype=torch.uint8, device=device)
            w[-256:].zero_()            # completion counters live in the tail (include/cgc_sm100.h, "workspace")
            self._work[key] = w
        return w.data_ptr(), self.work_bytes

    def _check(self, rc, name):
        self.launches += 1
        if rc != 0:
            msg = self.lib.cgc_last_error()
            raise CgcError('%s failed (%d): %s' % (name, rc, msg.decode() if msg else ''))

    # -- graph construction
    def count_nodes(self, batch, num_nodes):
        num_nodes.zero_()
        self._check(self.lib.cgc_count_nodes(_ptr(batch, torch.int64), batch.numel(), num_nodes.numel(),
                                             _ptr(num_nodes, torch.int32), self._stream()), 'cgc_count_nodes')

    def pad_nodes(self, x, batch, start, nmax, out):
        self._check(self.lib.cgc_pad_nodes(_ptr(x, torch.float32), _ptr(batch, torch.int64),
                                           _ptr(start, torch.int32), x.shape[0], x.shape[1], nmax,
                                           out.shape[0] // max(nmax, 1), _ptr(out, torch.float32),
                                           self._stream()), 'cgc_pad_nodes')

    def coo_to_csr_workspace(self, cap, rows):
        return int(self.lib.cgc_coo_to_csr_workspace(cap, rows))

    def coo_to_csr(self, edge_index, batch, start, num_nodes, nmax, num_graphs, n_total,
                   add_self_loops, rowptr, col, work):
        self._check(self.lib.cgc_coo_to_csr(_ptr(edge_index, torch.int64), edge_index.shape[1],
                                            _ptr(batch, torch.int64), _ptr(start, torch.int32), n_total,
                                            num_graphs, nmax, int(add_self_loops),
                                            _ptr(rowptr, torch.int32), _ptr(col, torch.int32),
                                            _ptr(work, torch.int32), work.numel(), self._stream()),
                    'cgc_coo_to_csr')

    def row_inv_degree(self, rowptr, rs):
        self._check(self.lib.cgc_row_inv_degree(_ptr(rowptr, torch.int32), rs.numel(),
                                                _ptr(rs, torch.float32), self._stream()), 'cgc_row_inv_degree')

    def edge_renorm(self, rowptr, col, p, val, rs):
        self._check(self.lib.cgc_edge_renorm(_ptr(rowptr, torch.int32), _ptr(col, torch.int32), rs.numel(),
                                             float(p), _ptr(val, torch.float32), _ptr(rs, torch.float32),
                                             self._stream()), 'cgc_edge_renorm')

    def csr_transpose(self, rowptr, col, val, colptr, row, val_t, work):
        self._check(self.lib.cgc_csr_transpose(_ptr(rowptr, torch.int32), _ptr(col, torch.int32),
                                               _ptr(val, torch.float32), rowptr.numel() - 1,
                                               _ptr(colptr, torch.int32), _ptr(row, torch.int32),
                                               _ptr(val_t, torch.float32), _ptr(work, torch.int32),
                                               work.numel(), self._stream()), 'cgc_csr_transpose')

    # -- cell-graph construction
    def radius_graph(self, pos, start, num_graphs, nmax, r, k, loop, nbr, deg):
        if not pos.is_contiguous() or pos.dim() != 2 or pos.shape[1] != 2:
            raise CgcError('cgc_radius_graph needs contiguous [n, 2] positions')
        self._check(self.lib.cgc_radius_graph(_ptr(pos, torch.float32), _ptr(start, torch.int32), num_graphs, nmax,
                                              float(r), int(k), int(bool(loop)), _ptr(nbr, torch.int32),
                                              _ptr(deg, torch.int32), self._stream()), 'cgc_radius_graph')

    def knn_to_coo(self, nbr, deg, offsets, k, num_edges, edge_index):
        self._check(self.lib.cgc_knn_to_coo(_ptr(nbr, torch.int32), _ptr(deg, torch.int32),
                                            _ptr(offsets, torch.int64), deg.numel(), int(k), int(num_edges),
                                            _ptr(edge_index, torch.int64), self._stream()), 'cgc_knn_to_coo')

    def farthest_point_sample(self, pos, start, first, out_start, nmax, out):
        if not pos.is_contiguous() or pos.dim() != 2 or pos.shape[1] != 2:
            raise CgcError('cgc_farthest_point_sample needs contiguous [n, 2] positions')
        self._check(self.lib.cgc_farthest_point_sample(_ptr(pos, torch.float32), _ptr(start, torch.int32),
                                                       _ptr(first, torch.int32), _ptr(out_start, torch.int32),
                                                       first.numel(), nmax, _ptr(out, torch.int32),
                                                       self._stream()), 'cgc_farthest_point_sample')

    # -- sparse gather
    def spmm(self, rowptr, col, val, x, src_scale, row_scale, out, accumulate):
        xp, ldx = _mat(x)
        op, ldo = _mat(out)
        self._check(self.lib.cgc_spmm(_ptr(rowptr, torch.int32), _ptr(col, torch.int32),
                                      _ptr(val, torch.float32), rowptr.numel() - 1, xp, ldx, x.shape[1],
                                      _ptr(src_scale, torch.float32), _ptr(row_scale, torch.float32),
                                      op, ldo, int(accumulate), self._stream()), 'cgc_spmm')

    def sage_gather_mean_fwd(self, rowptr, col, val, row_scale, x, num_graphs, nmax, out):
        if not (x.is_contiguous() and out.is_contiguous()):
            raise CgcError('cgc_sage_gather_mean_fwd needs contiguous x / out')
        self._check(self.lib.cgc_sage_gather_mean_fwd(_ptr(rowptr, torch.int32), _ptr(col, torch.int32),
                                                      _ptr(val, torch.float32),
                                                      _ptr(row_scale, torch.float32),
                                                      _ptr(x, torch.float32), num_graphs, nmax,
                                                      x.shape[1], _ptr(out, torch.float32),
                                                      self._stream()), 'cgc_sage_gather_mean_fwd')

    def ell16_rows(self, num_graphs, nmax):
        return int(self.lib.cgc_ell16_rows(num_graphs, nmax))

    def csr_to_ell16(self, rowptr, col, val, row_scale, num_graphs, nmax, width, pair_rows, ell, ell_val,
                     ell_scale, perm):
        rows = self.ell16_rows(num_graphs, nmax)
        if ell.dtype != torch.int16 or ell.numel() < rows * width or perm.numel() < rows // 8 \
                or ell_scale.numel() < rows:
            raise CgcError('cgc_csr_to_ell16: ell must hold ell16_rows * width int16, ell_scale ell16_rows '
                           'floats, perm ell16_rows / 8 int32')
        self._check(self.lib.cgc_csr_to_ell16(_ptr(rowptr, torch.int32), _ptr(col, torch.int32),
                                              _ptr(val, torch.float32), _ptr(row_scale, torch.float32),
                                              num_graphs, nmax, width, int(pair_rows), _ptr(ell, torch.int16),
                                              _ptr(ell_val, torch.float32), _ptr(ell_scale, torch.float32),
                                              _ptr(perm, torch.int32), self._stream()), 'cgc_csr_to_ell16')

    def sage_gather_ell_fwd(self, ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, num_graphs, nmax,
                            out):
        if not (x.is_contiguous() and out.is_contiguous()):
            raise CgcError('cgc_sage_gather_ell_fwd needs contiguous x / out')
        self._check(self.lib.cgc_sage_gather_ell_fwd(_ptr(ell, torch.int16), _ptr(ell_val, torch.float32),
                                                     _ptr(ell_scale, torch.float32), _ptr(perm, torch.int32),
                                                     width, _ptr(rowptr, torch.int32),
                                                     _ptr(col, torch.int32), _ptr(val, torch.float32),
                                                     _ptr(x, torch.float32), num_graphs, nmax, x.shape[1],
                                                     _ptr(out, torch.float32), self._stream()),
                    'cgc_sage_gather_ell_fwd')

    def sage_gather_ell(self, ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, num_graphs, nmax, out,
                        accumulate=False, tables_ready=False, pdl=True, two_per_sm=None, multicast=True):
        """General ELL16 gather: x / out may be column slices (leading dimensions); ``accumulate`` adds into out;
        ``tables_ready`` = the tables were not written by the kernel enqueued right before (PDL prologue);
        ``two_per_sm`` None = the library's choice, True / False force one kernel of the roofline configuration."""
        xp, ldx = _mat(x)
        op, ldo = _mat(out)
        flags = (GATHER_ACCUMULATE if accumulate else 0) | (GATHER_TABLES_READY if tables_ready else 0) | \
                (0 if pdl else GATHER_NO_PDL) | \
                (0 if two_per_sm is None else (GATHER_TWO_PER_SM if two_per_sm else GATHER_ONE_PER_SM)) | \
                (0 if multicast else GATHER_NO_MULTICAST)
        self._check(self.lib.cgc_sage_gather_ell(_ptr(ell, torch.int16), _ptr(ell_val, torch.float32),
                                                 _ptr(ell_scale, torch.float32), _ptr(perm, torch.int32), width,
                                                 _ptr(rowptr, torch.int32), _ptr(col, torch.int32),
                                                 _ptr(val, torch.float32), xp, ldx, num_graphs, nmax, x.shape[1],
                                                 op, ldo, flags, self._stream()), 'cgc_sage_gather_ell')

    def sage_layer_ell_supported(self, cin, cout, nmax):
        return bool(self.lib.cgc_sage_layer_ell_supported(int(cin), int(cout), int(nmax)))

    def sage_layer_ell_fwd(self, ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, num_graphs, nmax, w, b,
                           mask, act, sa, y, rmean, rvar, momentum, eps, mean, invstd, tables_ready=False, pdl=True):
        """Fused narrow SAGE layer (gather + W + bias + L2 norm + mask + act + BatchNorm statistics); ``mean`` None =
        no statistics (eval mode / no BatchNorm)."""
        xp, ldx = _mat(x)
        yp, ldy = _mat(y)
        if not sa.is_contiguous():
            raise CgcError('cgc_sage_layer_ell_fwd needs a contiguous sa')
        nn, _ = _mask(mask)
        wp, wb = self._workspace(x.device)
        self._check(self.lib.cgc_sage_layer_ell_fwd(_ptr(ell, torch.int16), _ptr(ell_val, torch.float32),
                                                    _ptr(ell_scale, torch.float32), _ptr(perm, torch.int32), width,
                                                    _ptr(rowptr, torch.int32), _ptr(col, torch.int32),
                                                    _ptr(val, torch.float32), xp, ldx, num_graphs, nmax, x.shape[1],
                                                    _ptr(w.contiguous(), torch.float32), _ptr(b, torch.float32), nn,
                                                    w.shape[1], int(act), _ptr(sa, torch.float32), yp, ldy,
                                                    _ptr(rmean, torch.float32), _ptr(rvar, torch.float32),
                                                    float(momentum), float(eps), _ptr(mean, torch.float32),
                                                    _ptr(invstd, torch.float32), wp, wb,
                                                    (GATHER_TABLES_READY if tables_ready else 0) |
                                                    (0 if pdl else GATHER_NO_PDL), self._stream()),
                    'cgc_sage_layer_ell_fwd')

    def gather_wide_ell(self, ell, ell_val, ell_scale, perm, width, rowptr, col, val, x, num_graphs, nmax, out,
                        accumulate):
        xp, ldx = _mat(x)
        op, ldo = _mat(out)
        self._check(self.lib.cgc_gather_wide_ell(_ptr(ell, torch.int16), _ptr(ell_val, torch.float32),
                                                 _ptr(ell_scale, torch.float32), _ptr(perm, torch.int32), width,
                                                 _ptr(rowptr, torch.int32), _ptr(col, torch.int32),
                                                 _ptr(val, torch.float32), xp, ldx, num_graphs, nmax, x.shape[1],
                                                 op, ldo, int(accumulate), self._stream()), 'cgc_gather_wide_ell')

    # -- SAGE epilogue
    def sage_epilogue_fwd(self, sa, w, b, mask, y, act=0):
        sp, lds = _mat(sa)
        yp, ldy = _mat(y)
        nn, nmax = _mask(mask)
        self._check(self.lib.cgc_sage_epilogue_fwd(sp, lds, _ptr(w.contiguous(), torch.float32),
                                                   _ptr(b, torch.float32), nn, nmax, sa.shape[0],
                                                   w.shape[0], w.shape[1], int(act), yp, ldy, self._stream()),
                    'cgc_sage_epilogue_fwd')

    def sage_layer_fwd_narrow(self, sa, w, b, mask, y, act, rmean, rvar, momentum, eps, mean, invstd):
        sp, lds = _mat(sa)
        yp, ldy = _mat(y)
        nn, nmax = _mask(mask)
        wp, wb = self._workspace(sa.device)
        self._check(self.lib.cgc_sage_layer_fwd_narrow(sp, lds, _ptr(w.contiguous(), torch.float32),
                                                       _ptr(b, torch.float32), nn, nmax, sa.shape[0], w.shape[0],
                                                       w.shape[1], int(act), yp, ldy, _ptr(rmean, torch.float32),
                                                       _ptr(rvar, torch.float32), float(momentum), float(eps),
                                                       _ptr(mean, torch.float32), _ptr(invstd, torch.float32),
                                                       wp, wb, self._stream()), 'cgc_sage_layer_fwd_narrow')

    def sage_layer_fwd_wide(self, sa, w, b, mask, y, act, rmean, rvar, momentum, eps, mean, invstd):
        sp, lds = _mat(sa)
        yp, ldy = _mat(y)
        nn, nmax = _mask(mask)
        wp, wb = self._workspace(sa.device)
        self._check(self.lib.cgc_sage_layer_fwd_wide(sp, lds, _ptr(w.contiguous(), torch.float32),
                                                     _ptr(b, torch.float32), nn, nmax, sa.shape[0], w.shape[0],
                                                     w.shape[1], int(act), yp, ldy, _ptr(rmean, torch.float32),
                                                     _ptr(rvar, torch.float32), float(momentum), float(eps),
                                                     _ptr(mean, torch.float32), _ptr(invstd, torch.float32),
                                                     wp, wb, self._stream()), 'cgc_sage_layer_fwd_wide')

    def sage_dh(self, gy, sa, w, b, y, mask, dh, act=0):
        gp, ldg = _mat(gy)
        sp, lds = _mat(sa)
        yp, ldy = _mat(y)
        dp, ldd = _mat(dh)
        nn, nmax = _mask(mask)
        self._check(self.lib.cgc_sage_dh(gp, ldg, sp, lds, _ptr(w.contiguous(), torch.float32),
                                         _ptr(b, torch.float32), yp, ldy, nn, nmax, sa.shape[0],
                                         w.shape[0], w.shape[1], int(act), dp, ldd, self._stream()), 'cgc_sage_dh')

    def sage_dh_bn(self, gx, sa, w, b, y, mean, invstd, gamma, d_gamma, d_beta, training, mask, dh, act=0):
        """sage_dh with the BatchNorm backward folded in: ``gx`` is the gradient w.r.t. the BatchNorm output."""
        gp, ldg = _mat(gx)
        sp, lds = _mat(sa)
        yp, ldy = _mat(y)
        dp, ldd = _mat(dh)
        nn, nmax = _mask(mask)
        self._check(self.lib.cgc_sage_dh_bn(gp, ldg, sp, lds, _ptr(w.contiguous(), torch.float32),
                                            _ptr(b, torch.float32), yp, ldy, _ptr(mean, torch.float32),
                                            _ptr(invstd, torch.float32), _ptr(gamma, torch.float32),
                                            _ptr(d_gamma, torch.float32), _ptr(d_beta, torch.float32), int(training),
                                            nn, nmax, sa.shape[0], w.shape[0], w.shape[1], int(act), dp, ldd,
                                            self._stream()), 'cgc_sage_dh_bn')

    def sage_layer_bwd_narrow(self, gy, sa, w, b, y, mask, dsa, d_w, d_b, act=0):
        gp, ldg = _mat(gy)
        sp, lds = _mat(sa)
        yp, ldy = _mat(y)
        dp, ldd = _mat(dsa)
        nn, nmax = _mask(mask)
        wp, wb = self._workspace(sa.device)
        self._check(self.lib.cgc_sage_layer_bwd_narrow(gp, ldg, sp, lds, _ptr(w.contiguous(), torch.float32),
                                                       _ptr(b, torch.float32), yp, ldy, nn, nmax, sa.shape[0],
                                                       w.shape[0], w.shape[1], int(act), dp, ldd,
                                                       _ptr(d_w, torch.float32), _ptr(d_b, torch.float32), wp, wb,
                                                       self._stream()), 'cgc_sage_layer_bwd_narrow')

    def act_fwd(self, x, act, y):
        xp, ldx = _mat(x)
        yp, ldy = _mat(y)
        self._check(self.lib.cgc_act_fwd(xp, ldx, x.shape[0], x.shape[1], int(act), yp, ldy, self._stream()),
                    'cgc_act_fwd')

    def act_bwd(self, gy, y, act, gx):
        gp, ldg = _mat(gy)
        yp, ldy = _mat(y)
        op, ldo = _mat(gx)
        self._check(self.lib.cgc_act_bwd(gp, ldg, yp, ldy, y.shape[0], y.shape[1], int(act), op, ldo,
                                         self._stream()), 'cgc_act_bwd')

    # -- DenseJK
    def jk_supported(self, channels, layers):
        return bool(self.lib.cgc_jk_supported(channels, layers))

    def jk_tape_floats(self, rows, channels, layers):
        return int(self.lib.cgc_jk_tape_floats(rows, channels, layers))

    def jk_lstm_fwd(self, feats, channels, layers, lstm_params, att_w, att_b, out, tape):
        fp, ld = _mat(feats)
        op, ldo = _mat(out)
        ptrs = [_ptr(p.contiguous(), torch.float32) for p in lstm_params]
        self._check(self.lib.cgc_jk_lstm_fwd(fp, ld, feats.shape[0], channels, layers, *ptrs,
                                             _ptr(att_w.contiguous(), torch.float32), _ptr(att_b, torch.float32),
                                             op, ldo, _ptr(tape, torch.float32), self._stream()),
                    'cgc_jk_lstm_fwd')

    def jk_lstm_bwd(self, g_out, feats, channels, layers, lstm_params, att_w, tape, g_feats, dgates, dscore):
        gp, ldg = _mat(g_out)
        fp, ld = _mat(feats)
        op, ldo = _mat(g_feats)
        w = [_ptr(lstm_params[i].contiguous(), torch.float32) for i in (0, 1, 4, 5)]
        self._check(self.lib.cgc_jk_lstm_bwd(gp, ldg, fp, ld, feats.shape[0], channels, layers, *w,
                                             _ptr(att_w.contiguous(), torch.float32), _ptr(tape, torch.float32),
                                             op, ldo, _ptr(dgates, torch.float32), _ptr(dscore, torch.float32),
                                             self._stream()), 'cgc_jk_lstm_bwd')

    def scale_rows_bwd(self, dsa, sa, rs, ga, d_rs):
        for t in (dsa, sa):
            if not t.is_contiguous():
                raise CgcError('cgc_scale_rows_bwd needs contiguous dsa / sa')
        gp, ldg = _mat(ga)
        self._check(self.lib.cgc_scale_rows_bwd(_ptr(dsa, torch.float32), _ptr(sa, torch.float32),
                                                _ptr(rs, torch.float32), dsa.shape[0], dsa.shape[1],
                                                gp, ldg, _ptr(d_rs, torch.float32),
                                                self._stream()), 'cgc_scale_rows_bwd')

    # -- batch norm
    def bn_stats(self, y, rmean, rvar, momentum, eps, mean, invstd):
        yp, ldy = _mat(y)
        wp, wb = self._workspace(y.device)
        self._check(self.lib.cgc_bn_stats(yp, ldy, y.shape[0], y.shape[1], _ptr(rmean, torch.float32),
                                          _ptr(rvar, torch.float32), float(momentum), float(eps),
                                          _ptr(mean, torch.float32), _ptr(invstd, torch.float32), wp, wb,
                                          self._stream()), 'cgc_bn_stats')

    def bn_eval_stats(self, rmean, rvar, eps, mean, invstd):
        self._check(self.lib.cgc_bn_eval_stats(_ptr(rmean, torch.float32), _ptr(rvar, torch.float32),
                                               rmean.numel(), float(eps), _ptr(mean, torch.float32),
                                               _ptr(invstd, torch.float32), self._stream()),
                    'cgc_bn_eval_stats')

    def bn_apply(self, y, mean, invstd, gamma, beta, mask, out):
        yp, ldy = _mat(y)
        op, ldo = _mat(out)
        nn, nmax = _mask(mask)
        self._check(self.lib.cgc_bn_apply(yp, ldy, _ptr(mean, torch.float32), _ptr(invstd, torch.float32),
                                          _ptr(gamma, torch.float32), _ptr(beta, torch.float32), nn, nmax,
                                          y.shape[0], y.shape[1], op, ldo, self._stream()), 'cgc_bn_apply')

    def bn_bwd(self, gx, y, mean, invstd, gamma, mask, training, gy, d_gamma, d_beta):
        gp, ldg = _mat(gx)
        yp, ldy = _mat(y)
        op, ldo = _mat(gy) if gy is not None else (None, 0)       # gy None: column sums only
        nn, nmax = _mask(mask)
        wp, wb = self._workspace(y.device)
        self._check(self.lib.cgc_bn_bwd(gp, ldg, yp, ldy, _ptr(mean, torch.float32),
                                        _ptr(invstd, torch.float32), _ptr(gamma, torch.float32), nn, nmax,
                                        y.shape[0], y.shape[1], int(training), op, ldo,
                                        _ptr(d_gamma, torch.float32), _ptr(d_beta, torch.float32), wp, wb,
                                        self._stream()), 'cgc_bn_bwd')

    # -- softmax / readout
    def softmax_fwd(self, z, mask, s, s_full, arg):
        nn, nmax = _mask(mask)
        if not z.is_contiguous() or (s_full is not None and not s_full.is_contiguous()):
            raise CgcError('cgc_softmax_fwd needs contiguous z / s_full')
        sp, lds = _mat(s)
        self._check(self.lib.cgc_softmax_fwd(_ptr(z, torch.float32), z.shape[0], z.shape[1], nn, nmax,
                                             sp, lds, _ptr(s_full, torch.float32),
                                             _ptr(arg, torch.int32), self._stream()), 'cgc_softmax_fwd')

    def softmax_bwd(self, gs, s, gz):
        gp, ldg = _mat(gs)
        sp, lds = _mat(s)
        zp, ldz = _mat(gz)
        if ldg != lds:
            raise CgcError('cgc_softmax_bwd: gs and s must share their leading dimension')
        self._check(self.lib.cgc_softmax_bwd(gp, sp, lds, s.shape[0], s.shape[1], zp, ldz,
                                             self._stream()), 'cgc_softmax_bwd')

    def colmax_fwd(self, x, b, n, out, arg):
        self._check(self.lib.cgc_colmax_fwd(_ptr(x, torch.float32), b, n, x.shape[1],
                                            _ptr(out, torch.float32), _ptr(arg, torch.int32),
                                            self._stream()), 'cgc_colmax_fwd')

    def colmax_bwd(self, g, arg, b, n, gx):
        self._check(self.lib.cgc_colmax_bwd(_ptr(g, torch.float32), _ptr(arg, torch.int32), b, n,
                                            g.shape[1], _ptr(gx, torch.float32), self._stream()),
                    'cgc_colmax_bwd')

    # -- dense adjacency helpers
    def row_inv_clamped_sum(self, adj, rs, rsum):
        adj = adj.contiguous()
        self._check(self.lib.cgc_row_inv_clamped_sum(_ptr(adj, torch.float32), rs.numel(), adj.shape[-1],
                                                     _ptr(rs, torch.float32), _ptr(rsum, torch.float32),
                                                     self._stream()), 'cgc_row_inv_clamped_sum')

    def add_rs_grad(self, d_adj, rsum, d_rs):
        self._check(self.lib.cgc_add_rs_grad(_ptr(d_adj, torch.float32), _ptr(rsum, torch.float32),
                                             _ptr(d_rs, torch.float32), rsum.numel(), d_adj.shape[-1],
                                             self._stream()), 'cgc_add_rs_grad')

    def dense_renorm_fwd(self, adj, p, mask, out, rho):
        nn, _ = _mask(mask)
        self._check(self.lib.cgc_dense_renorm_fwd(_ptr(adj, torch.float32), adj.shape[0], adj.shape[1],
                                                  float(p), nn, _ptr(out, torch.float32),
                                                  _ptr(rho, torch.float32), self._stream()),
                    'cgc_dense_renorm_fwd')

    def dense_renorm_bwd(self, g, out, rho, p, mask, g_adj):
        nn, _ = _mask(mask)
        self._check(self.lib.cgc_dense_renorm_bwd(_ptr(g, torch.float32), _ptr(out, torch.float32),
                                                  _ptr(rho, torch.float32), g.shape[0], g.shape[1],
                                                  float(p), nn, _ptr(g_adj, torch.float32),
                                                  self._stream()), 'cgc_dense_renorm_bwd')

    # -- dense contractions
    def gemm(self, a, b, c, ta, tb, alpha=1.0, beta=0.0, bias=None, row_scale=None, rowmask=None,
             precision=None):
        def geom(t):
            if t.dim() == 2:
                if t.shape[1] > 1 and t.stride(1) != 1:
                    raise CgcError('gemm operand must be row-major')
                return 1, t.shape[0], t.shape[1], t.stride(0), 0
            if t.dim() != 3 or (t.shape[2] > 1 and t.stride(2) != 1):
                raise CgcError('gemm operand must be [batch, rows, cols] row-major')
            return t.shape[0], t.shape[1], t.shape[2], t.stride(1), t.stride(0)
        ba, ra, ca, lda, sa_ = geom(a)
        bb, rb, cb, ldb, sb_ = geom(b)
        bc, m, n, ldc, sc_ = geom(c)
        k = ra if ta else ca
        if (ca if ta else ra) != m or (rb if tb else cb) != n or (cb if tb else rb) != k \
                or ba != bc or bb != bc:
            raise CgcError('gemm shape mismatch: a%s b%s c%s ta=%s tb=%s'
                           % (tuple(a.shape), tuple(b.shape), tuple(c.shape), ta, tb))
        nn, nmax = _mask(rowmask)
        wp, wb = self._workspace(c.device)
        prec = self.gemm_precision if precision is None else precision
        self._check(self.lib.cgc_gemm(int(ta), int(tb), m, n, k, float(alpha),
                                      _ptr(a, torch.float32), lda, sa_, _ptr(b, torch.float32), ldb, sb_,
                                      float(beta), _ptr(c, torch.float32), ldc, sc_, bc,
                                      _ptr(bias, torch.float32), _ptr(row_scale, torch.float32), nn, nmax,
                                      prec, wp, wb, self._stream()), 'cgc_gemm')

    def colsum(self, x, mask, out):
        xp, ldx = _mat(x)
        nn, nmax = _mask(mask)
        wp, wb = self._workspace(x.device)
        self._check(self.lib.cgc_colsum(xp, ldx, x.shape[0], x.shape[1], nn, nmax,
                                        _ptr(out, torch.float32), wp, wb, self._stream()), 'cgc_colsum')


_active = None


def get():
    """The backend of the product path: ``CudaBackend`` (raises if the library is missing)."""
    global _active
    if _active is None:
        _active = CudaBackend()
    return _active


def use_backend_for_testing(backend):
    """TESTS ONLY: substitute the backend (``None`` restores the CUDA library)."""
    global _active
    _active = backend
