"""CPU: the C-ABI shared library loads and exports exactly what include/mvd_b200.h declares; host-only entry points
and argument validation behave (no kernel is launched here)."""
import ctypes
import os
import re

import torch

from mvdfusion_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mvd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mvd_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared  # the ctypes table binds every entry point, no more, no less
    assert lib.mvd_abi_version() == _lib.ABI_VERSION


def test_geglu_permutation_matches_emulation():
    from ops_double import geglu_permutation
    lib = _lib.load()
    for inner, tile in ((1280, 128), (256, 64), (5120, 256)):
        buf = (ctypes.c_int32 * (2 * inner))()
        assert lib.mvd_geglu_row_permutation(inner, tile, ctypes.cast(buf, ctypes.c_void_p)) == 0
        perm = torch.tensor(list(buf))
        assert torch.equal(perm, geglu_permutation(inner, tile))
        assert torch.equal(torch.sort(perm).values, torch.arange(2 * inner))  # a permutation


def test_argument_validation_reports_errors():
    lib = _lib.load()
    assert lib.mvd_gemm_f16(None, None) == -1
    assert b"null" in lib.mvd_last_error()
    g = _lib.GemmArgs()
    g.M, g.N, g.K = 128, 128, 0
    assert lib.mvd_gemm_f16(ctypes.byref(g), None) == -1
    assert lib.mvd_attn_self_f16(None, None, None, None, 1, 8, 1024, 40, 64, 320, None) == -1
    assert lib.mvd_geglu_row_permutation(100, 100, None) == -1
    assert lib.mvd_groupnorm_f32_f16(None, None, None, None, None, 1, 1, 32, 1e-5, 0, None) == -1
    # ABI 15-17 (training kernels): null pointers, channel counts, alignment
    assert lib.mvd_layernorm_fwd_f32(None, None, None, None, None, 4, 320, 1e-5, None) == -1
    assert lib.mvd_layernorm_bwd_f32(None, None, None, None, None, None, None, 4, 320, None) == -1
    assert lib.mvd_groupnorm_fwd_f32(None, None, None, None, None, None, 1, 16, 320, 1e-5, 1, None) == -1
    assert lib.mvd_groupnorm_bwd_f32(None, None, None, None, None, None, None, None, None, 1, 16, 320, 1, None) == -1
    assert lib.mvd_act_fwd_f32(None, None, 4, 4, 1, None) == -1
    assert lib.mvd_bilinear_gather_fwd_f32(None, None, None, 1, 8, 8, 256, 10, None) == -1
    buf = (ctypes.c_float * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.mvd_layernorm_fwd_f32(p, None, None, p, p, 4, 322, 1e-5, None) == -1          # C % 4
    assert lib.mvd_layernorm_fwd_f32(p, p, None, p, p, 4, 320, 1e-5, None) == -1            # gamma without beta
    assert lib.mvd_groupnorm_fwd_f32(p, p, p, p, p, p, 1, 4, 48, 1e-5, 0, None) == -1       # C % 32
    assert lib.mvd_act_fwd_f32(p, p, 4, 4, 7, None) == -1 and b"mode" in lib.mvd_last_error()
    assert lib.mvd_act_fwd_f32(p, p, 4, 6, 3, None) == -3                                    # GEGLU needs cols % 4 == 0
    assert lib.mvd_bilinear_gather_fwd_f32(p, p, p, 1, 8, 8, 6, 10, None) == -1              # C % 4
    assert lib.mvd_launch_count() == 0  # validation failures launch nothing


def test_product_refuses_cpu_tensors():
    import pytest
    from mvdfusion_b200.ops import MvdError
    from mvdfusion_b200.runtime import get_ops
    with pytest.raises(MvdError):
        get_ops("cpu")
