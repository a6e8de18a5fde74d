"""The C-ABI library loads and exports exactly what include/cgc_sm100.h declares (no compute)."""
import os
import re

import pytest

from cgc_net_b200 import _backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'cgc_sm100.h')).read()
    return set(re.findall(r'CGC_API\s+[\w\s\*]+?\b(cgc_\w+)\s*\(', text))


def test_header_and_binding_agree():
    assert _declared() == set(_backend.EXPORTED_SYMBOLS)


def test_library_exports_every_symbol():
    if not os.path.isfile(_backend.LIB_PATH):
        pytest.fail('libcgc_sm100.so is not built: run __graft_entry__.build()')
    lib = _backend.load_library()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.cgc_version() >= 100
    assert lib.cgc_workspace_bytes() > 0
    assert lib.cgc_coo_to_csr_workspace(1000, 100) > 1000


def test_argument_checks_need_no_gpu():
    """Entry points validate arguments before touching the device and report through cgc_last_error."""
    lib = _backend.load_library()
    rc = lib.cgc_row_inv_degree(None, 0, None, None)
    assert rc == -1
    assert b'cgc_row_inv_degree' in lib.cgc_last_error()
    rc = lib.cgc_gemm(0, 0, 0, 4, 4, 1.0, None, 4, 0, None, 4, 0, 0.0, None, 4, 0, 1, None, None, None, 0, -1,
                      None, 0, None)
    assert rc == -1
