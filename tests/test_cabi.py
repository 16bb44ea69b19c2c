"""The C-ABI shared library loads on a machine without a GPU and exports every symbol that
include/sympa_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from sympa_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build_library()
    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "sympa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sympa_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_lists():
    from sympa_b200 import _lib
    assert declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_host_only_entry_points(lib):
    assert lib.sympa_version() == 4
    assert lib.sympa_error_string(0) == b"ok"
    assert lib.sympa_error_string(2).startswith(b"unsupported")
    # 2 operands * pairs * (2 n n) doubles * 8 bytes
    # saved state: packed lower triangles of the unit gradients, for every kernel family (ABI 4)
    assert lib.sympa_workspace_bytes(0, 4, 1000) == 2 * 1000 * 20 * 8
    assert lib.sympa_workspace_bytes(2, 3, 10) == 2 * 10 * 6 * 8
    assert lib.sympa_workspace_bytes(0, 10, 10) == 2 * 10 * 110 * 8
    # packed gradient table of the table backward: one packed point per row
    assert lib.sympa_backward_workspace_bytes(0, 10, 1000) == 1000 * 110 * 8
    assert lib.sympa_backward_workspace_bytes(2, 7, 1000) == 1000 * 28 * 8
    assert lib.sympa_workspace_bytes(0, 11, 10) == -1
    # scratch: only upper n > 4 and batches worth splitting; five n x n planes + n per pair of a chunk, + (1 + n) per pair
    assert lib.sympa_scratch_bytes(0, 10, 1 << 20) == 0        # split path is off by default
    assert lib.sympa_set_option(1, 1) == 0
    try:
        assert lib.sympa_scratch_bytes(0, 4, 1 << 20) == 0
        assert lib.sympa_scratch_bytes(1, 10, 1 << 20) == 0
        assert lib.sympa_scratch_bytes(0, 10, 100) == 0
        assert lib.sympa_scratch_bytes(0, 10, 1 << 20) == 32768 * 510 * 8 + (1 << 20) * 11 * 8
    finally:
        assert lib.sympa_set_option(1, 0) == 0
    assert lib.sympa_set_option(99, 0) == 1
    assert lib.sympa_set_option(2, 0) == 1 and lib.sympa_set_option(2, 40) == 0     # scatter pass size in MB


def test_argument_errors_are_synchronous(lib):
    P = ctypes.c_void_p
    # unsupported n, and neither / both operand forms
    assert lib.sympa_dist_forward(0, 64, 0, 1, P(8), P(8), None, 0, None, None, P(8), None, None, None, 0, None, None) == 2
    assert lib.sympa_dist_forward(0, 2, 0, 1, None, None, None, 0, None, None, P(8), None, None, None, 0, None, None) == 1
    assert lib.sympa_dist_forward(0, 2, 0, 1, P(8), P(8), P(8), 4, P(8), None, P(8), None, None, None, 0, None, None) == 1
    assert lib.sympa_dist_forward(0, 2, 4, 1, P(8), P(8), None, 0, None, None, P(8), None, None, None, 0, None, None) == 1  # wsum without weights
    assert lib.sympa_dist_forward(0, 2, 0, 0, P(8), P(8), None, 0, None, None, P(8), None, None, None, 0, None, None) == 0  # empty batch


def test_ops_refuse_cpu_tensors():
    import torch
    from sympa_b200 import UpperHalfManifold
    m = UpperHalfManifold(dims=2)
    x = m.random(3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.dist(x, x)
