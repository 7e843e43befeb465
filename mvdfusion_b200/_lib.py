"""ctypes binding of libmvd_b200.so (the C ABI declared in include/mvd_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  `load()` raises if the shared object is
missing or does not export every symbol the header declares.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# MVD_B200_LIB: another build of the same library (e.g. the instrumented `make trace` one); it must exist — there is no fallback
LIB_PATH = os.environ.get("MVD_B200_LIB") or os.path.join(_HERE, "libmvd_b200.so")
ABI_VERSION = 17


class GemmArgs(ctypes.Structure):
    """struct mvd_gemm_args (include/mvd_b200.h)"""

    _fields_ = [
        ("M", c_int32), ("N", c_int32), ("K", c_int32), ("a_mode", c_int32),
        ("A", c_void_p), ("lda", c_int32),
        ("n_img", c_int32), ("H", c_int32), ("W", c_int32), ("C", c_int32),
        ("Wt", c_void_p), ("ldw", c_int32),
        ("bias", c_void_p), ("rowbias", c_void_p), ("rows_per_group", c_int32),
        ("colscale", c_void_p), ("residual", c_void_p), ("ldr", c_int32),
        ("act", c_int32), ("out_mode", c_int32),
        ("out", c_void_p), ("ldc", c_int32),
        ("out_k", c_void_p), ("out_vt", c_void_p),
        ("heads", c_int32), ("dhead", c_int32), ("dpad", c_int32), ("seq", c_int32),
        ("split_k", c_int32), ("tile_n", c_int32), ("cta_pair", c_int32),
        ("splitk_ws", c_void_p), ("splitk_ws_bytes", c_longlong),
        ("out16", c_void_p), ("ld16", c_int32), ("hilo", c_int32), ("out16_lo", c_int32), ("a_lo_off", c_int32), ("conv_stride", c_int32), ("conv_no_pad_lo", c_int32),
        ("ln_stats_out", c_void_p), ("ln_stats", c_void_p), ("ln_colsum", c_void_p), ("ln_eps", c_float), ("conv_up2", c_int32),
    ]


class DitLayer(ctypes.Structure):
    """struct mvd_dit_layer (include/mvd_b200.h)"""

    _fields_ = [(n, c_void_p) for n in ("w_qkv", "b_qkv", "w_proj", "b_proj", "w_fc1", "b_fc1", "w_fc2", "b_fc2",
                                        "shift_msa", "scale_msa", "shift_mlp", "scale_mlp")]


class DitArgs(ctypes.Structure):
    """struct mvd_dit_args (include/mvd_b200.h)"""

    _fields_ = [
        ("R", c_int32), ("V", c_int32), ("layers", c_int32), ("token_k", c_int32), ("token_ld", c_int32),
        ("tokens", c_void_p), ("w_pre", c_void_p), ("w_pre_ld", c_int32), ("b_pre", c_void_p),
        ("layer", DitLayer * 4),
        ("pool_w", c_void_p), ("pool_b", c_void_p), ("pooled", c_void_p), ("x_out", c_void_p), ("eps", c_float),
    ]


class FoldJob(ctypes.Structure):
    """struct mvd_fold_job (include/mvd_b200.h)"""

    _fields_ = [("w", c_void_p), ("gate", c_void_p), ("bias", c_void_p), ("w_out", c_void_p), ("b_out", c_void_p), ("N", c_int32), ("K", c_int32)]


i32, vp, f32, i64 = c_int32, c_void_p, c_float, c_longlong

# name -> argtypes (every entry returns int unless listed in _RESTYPE)
SIGNATURES = {
    "mvd_last_error": [],
    "mvd_abi_version": [],
    "mvd_launch_count": [],
    "mvd_gemm_f16": [ctypes.POINTER(GemmArgs), vp],
    "mvd_geglu_row_permutation": [i32, i32, vp],
    "mvd_gridattn_dit_f16": [ctypes.POINTER(DitArgs), vp],
    "mvd_dit_fold_gates": [ctypes.POINTER(FoldJob), i32, vp],
    "mvd_attn_self_f16": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "mvd_attn_self_masked_f16": [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "mvd_groupnorm_f32_f16": [vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, vp],
    "mvd_groupnorm_hilo_f32_f16": [vp, vp, vp, vp, i32, i32, i32, f32, i32, vp],
    "mvd_groupnorm2_f32_f16": [vp, i32, vp, i32, vp, vp, vp, i32, i32, f32, i32, vp],
    "mvd_layernorm_f32_f16": [vp, vp, vp, vp, i32, i32, f32, vp],
    "mvd_layernorm_f32_f32": [vp, i64, vp, vp, vp, i64, i32, i32, f32, vp],
    "mvd_ln_modulate_f32_f16": [vp, vp, vp, vp, i32, i32, f32, vp],
    "mvd_softmax_rows_f32_f16": [vp, vp, i32, i32, i32, i32, f32, vp],
    "mvd_cast_f32_f16": [vp, vp, i64, vp],
    "mvd_concat_f32": [vp, vp, vp, i64, i32, i32, vp],
    "mvd_concat_f32_f16": [vp, vp, vp, i64, i32, i32, vp],
    "mvd_upsample2x_f32_f16": [vp, vp, i32, i32, i32, i32, vp],
    "mvd_im2col_s2_f32_f16": [vp, vp, i32, i32, i32, i32, vp],
    "mvd_im2col_s2_pad_f32_f16": [vp, vp, i32, i32, i32, i32, i32, vp],
    "mvd_gemv_f16": [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "mvd_gemv_grouped_f16": [vp, i32, i32, vp, i32, i32, vp],
    "mvd_timestep_embedding": [vp, vp, vp, i32, vp],
    "mvd_unet_input_f16": [vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, vp],
    "mvd_cfg_ddim": [vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp],
    "mvd_nchw_to_rows_f32": [vp, vp, i32, i32, i32, vp],
    "mvd_rows_to_nchw_f32": [vp, vp, i32, i32, i32, i32, vp],
    "mvd_nchw_to_nhwc_f16": [vp, vp, i32, i32, i32, i32, vp],
    "mvd_gather_rows_f32": [vp, i64, vp, vp, vp],
    "mvd_increment_i32": [vp, i32, vp],
    "mvd_gridattn_prep": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, vp],
    "mvd_gridattn_tokens": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "mvd_view_attention_f16": [vp, vp, i32, i32, i32, i32, vp],
    "mvd_view_pool_f16": [vp, vp, vp, vp, i32, i32, i32, vp],
    "mvd_frustum_pool_f16": [vp, vp, i32, i32, i32, i32, i32, vp],
    "mvd_pixel_cross_attn_f16": [vp, vp, vp, i32, i32, i32, i32, vp],
    "mvd_layernorm_fwd_f32": [vp, vp, vp, vp, vp, i32, i32, f32, vp],
    "mvd_layernorm_bwd_f32": [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp],
    "mvd_groupnorm_fwd_f32": [vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, vp],
    "mvd_groupnorm_bwd_f32": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "mvd_act_fwd_f32": [vp, vp, i64, i32, i32, vp],
    "mvd_act_bwd_f32": [vp, vp, vp, i64, i32, i32, vp],
    "mvd_bilinear_gather_fwd_f32": [vp, vp, vp, i32, i32, i32, i32, i64, vp],
    "mvd_bilinear_gather_bwd_f32": [vp, vp, vp, i32, i32, i32, i32, i64, vp],
}
_RESTYPE = {"mvd_last_error": c_char_p, "mvd_launch_count": c_longlong}

_lib = None


def load():
    """dlopen libmvd_b200.so and type every entry point.  Raises (never falls back) when unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C mvdfusion_b200/csrc` (or __graft_entry__.build()). "
            "mvdfusion_b200 has no CPU / PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RuntimeError(f"libmvd_b200.so does not export {name}; rebuild it") from e
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, c_int32)
    ver = lib.mvd_abi_version()
    if ver != ABI_VERSION:
        raise RuntimeError(f"libmvd_b200.so ABI version {ver} != expected {ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def last_error():
    return load().mvd_last_error().decode("utf-8", "replace")


def launch_count():
    return int(load().mvd_launch_count())
