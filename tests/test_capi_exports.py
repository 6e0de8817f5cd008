"""CPU: the C-ABI library loads and exports every symbol include/recmv_b200.h declares (no compute
calls without a GPU), and the host-side error behaviour mirrors the reference's."""
import ctypes

import pytest
import torch

from recmv_b200 import _lib, ops


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    declared = set(_lib.header_symbols())
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.recmv_version() >= 100
    assert lib.recmv_error_string(0) == b"ok"
    assert b"NULL" in lib.recmv_error_string(-1)
    assert lib.recmv_sdf_packed_bytes() > 1966592 * 4  # fp32 copy + fp16 hi/lo planes


def test_argument_errors_without_gpu():
    lib = _lib.load()
    # argument validation happens before any CUDA call
    assert lib.recmv_minv3x3_fwd(None, None, None, 5, 0, None) == -1
    assert lib.recmv_minv3x3_fwd(None, None, None, -1, 0, None) == -3
    assert lib.recmv_minv3x3_fwd(None, None, None, 0, 0, None) == 0
    nbytes = ctypes.c_size_t(0)
    assert lib.recmv_mc_scratch_bytes(257, 257, 257, ctypes.byref(nbytes)) == 0
    assert nbytes.value >= 257 ** 3 * 4 + 257 ** 3 // 8     # packed vertex words + 1-bit sign mask
    assert lib.recmv_mc_scratch_bytes(0, 4, 4, ctypes.byref(nbytes)) == -3
    with pytest.raises(_lib.RecmvError):
        _lib.check(-2, "x")


def test_host_shims_reject_cpu_tensors_like_the_reference():
    # CHECK_INPUT semantics (FastMinv/M3x3Inv.cpp:4-6): CPU tensor -> RuntimeError, never a fallback
    with pytest.raises(RuntimeError):
        ops.minv3x3(torch.randn(4, 3, 3))
    with pytest.raises(RuntimeError):
        ops.mc_gpu(torch.randn(4, 4, 4))
    from recmv_b200.model import getTmpSdf
    net = getTmpSdf("cpu", 6)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            net(torch.randn(8, 3), None)
