"""Tensor-level view of the C ABI: every method validates its torch tensors, freezes the raw pointers and sizes
into a bound call and returns it; `call(stream)` launches the kernel(s) asynchronously on that CUDA stream.

A step of the denoising loop is a fixed list of such bound calls (engine.Program), replayed either directly
or from a captured CUDA graph.  There is exactly one implementation of this interface in the product — the
sm_100a library — and it refuses anything that is not a CUDA tensor.
"""
import ctypes

import torch

from . import _lib

A_ROWMAJOR, A_CONV3X3 = 0, 1
ACT_NONE, ACT_GELU, ACT_SILU, ACT_GEGLU = 0, 1, 2, 3
GN_WS_BYTES_PER_IMAGE = 16384   # mvd_groupnorm_{fwd,bwd}_f32 scratch per image: <= 32 chunks x 32 groups x 2 fp64 partials
OUT_F32, OUT_F16, OUT_QKV_HEADS = 0, 1, 2


class MvdError(RuntimeError):
    pass


def gemm_signature(is_conv, M, N, K, out_kind, has_res, act, extra=""):
    """Key of a GEMM launch in mvdfusion_b200/gemm_tuning.json (tools/tune_gemm.py).  extra: "+st" for a producer that also leaves the
    fp16 copy and LayerNorm statistics of its rows, "+ln" for a consumer with the LayerNorm folded in (their epilogues cost differently)."""
    return f"{'conv' if is_conv else 'lin'}:{M}:{N}:{K}:{out_kind}:{int(bool(has_res))}:{act}{extra}"


def _ptr(t, dtype=None):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise MvdError("mvdfusion_b200 kernels take CUDA tensors only (no CPU fallback exists)")
    if dtype is not None and t.dtype != dtype:
        raise MvdError(f"expected {dtype}, got {t.dtype}")
    return t.data_ptr()


class BoundCall:
    """One C-ABI call with frozen arguments.  Keeps the tensors alive that its pointers refer to."""

    __slots__ = ("fn", "args", "keep", "name", "meta")

    def __init__(self, name, fn, args, keep, meta=None):
        self.name, self.fn, self.args, self.keep = name, fn, args, keep
        self.meta = meta or {}  # algorithmic flops / bytes of the launch (bench.py roofline accounting)

    def __call__(self, stream):
        rc = self.fn(*self.args, stream)
        if rc != 0:
            raise MvdError(f"{self.name} failed (rc={rc}): {_lib.last_error()}")


class NativeOps:
    """The sm_100a kernels of libmvd_b200.so (include/mvd_b200.h)."""

    def __init__(self, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise MvdError("NativeOps needs a CUDA device; mvdfusion_b200 has no CPU path")

    # ------------------------------------------------------------------ helpers
    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def geglu_permutation(self, inner, tile_n):
        perm = (ctypes.c_int32 * (2 * inner))()
        rc = self.lib.mvd_geglu_row_permutation(inner, tile_n, ctypes.cast(perm, ctypes.c_void_p))
        if rc != 0:
            raise MvdError(_lib.last_error())
        return torch.tensor(list(perm), dtype=torch.long)

    def _bind(self, name, args, keep, meta=None):
        return BoundCall(name, getattr(self.lib, name), tuple(args), keep, meta)

    # ------------------------------------------------------------------ GEMM / conv
    def gemm(self, A, Wt, out, M, N, K, *, lda=None, ldw=None, ldc=None, bias=None, rowbias=None, rows_per_group=1,
             colscale=None, residual=None, ldr=0, act=ACT_NONE, conv=None, qkv=None, split_k=1, tile_n=0, ws=None, cta_pair=0,
             out16=None, ld16=None, hilo=False, out16_lo=0, a_lo_off=0, conv_stride=1, conv_no_pad_lo=False,
             ln_stats_out=None, ln=None, conv_up2=False):
        """conv_up2: nearest x2 upsample folded into the convolution (ABI 14) — conv = (n_img, H, W, C) is the SOURCE image, M = n_img*H*W,
        N = 4 * Cout (phase blocks), K = 4 * C, `out` / `out16` have the upsampled resolution (4 M rows, Cout columns).
        ln_stats_out: fp32 [N/32, M, 2] — the epilogue leaves (sum, sum of squares) per row and 32-column chunk of an fp32 output;
        ln = (stats, colsum, eps): nn.LayerNorm folded into this GEMM — A holds the raw rows, Wt the gamma-scaled weights, stats what the
        GEMM that produced A's rows wrote (include/mvd_b200.h, ABI 13).
        out16: optional fp16 tensor (or column window of a wider one, row pitch ld16) that receives a copy of an fp32 output
        (out16_lo > 0: and its fp16 rounding residual, out16_lo columns to the right).
        hilo: split-precision operands A = [A_hi | A_lo], Wt = [W_hi | W_lo] (include/mvd_b200.h, ABI 9); K stays the logical K.
        conv = (n_img, H, W, C) with conv_stride = 2: H, W are the OUTPUT extent and A is the [n_img, 2H, 2W, C] image (ABI 10);
        in conv mode `lda` is the pixel pitch of A when it is a column window of a wider buffer."""
        g = _lib.GemmArgs()
        g.M, g.N, g.K = M, N, K
        g.A = _ptr(A, torch.float16)
        g.Wt = _ptr(Wt, torch.float16)
        g.ldw = ldw if ldw is not None else Wt.shape[-1]
        g.hilo = int(bool(hilo))
        g.out16_lo = int(out16_lo)
        g.a_lo_off = int(a_lo_off)
        if conv is not None:
            g.a_mode = A_CONV3X3
            g.n_img, g.H, g.W, g.C = conv
            g.lda = lda if lda is not None else 0
            g.conv_stride = conv_stride
            g.conv_no_pad_lo = int(bool(conv_no_pad_lo))
            g.conv_up2 = int(bool(conv_up2))
        else:
            g.a_mode = A_ROWMAJOR
            g.lda = lda if lda is not None else A.shape[-1]
        g.bias = _ptr(bias, torch.float32)
        g.rowbias = _ptr(rowbias, torch.float32)
        g.rows_per_group = rows_per_group
        g.colscale = _ptr(colscale, torch.float32)
        g.residual = _ptr(residual, torch.float32)
        g.ldr = ldr
        g.act = act
        g.out = _ptr(out)
        if qkv is not None:
            g.out_mode = OUT_QKV_HEADS
            g.out_k = _ptr(qkv["out_k"], torch.float16)
            g.out_vt = _ptr(qkv["out_vt"], torch.float16)
            g.heads, g.dhead, g.dpad, g.seq = qkv["heads"], qkv["dhead"], qkv["dpad"], qkv["seq"]
        else:
            g.out_mode = OUT_F16 if out.dtype == torch.float16 else OUT_F32
            g.ldc = ldc if ldc is not None else out.shape[-1]
        g.split_k = split_k
        g.tile_n = tile_n
        g.cta_pair = cta_pair
        if ws is not None:
            g.splitk_ws = _ptr(ws, torch.uint8)
            g.splitk_ws_bytes = ws.numel()
        if out16 is not None:
            g.out16 = _ptr(out16, torch.float16)
            g.ld16 = ld16 if ld16 is not None else out16.shape[-1]
        if ln_stats_out is not None:
            g.ln_stats_out = _ptr(ln_stats_out, torch.float32)
        if ln is not None:
            g.ln_stats = _ptr(ln[0], torch.float32)
            g.ln_colsum = _ptr(ln[1], torch.float32)
            g.ln_eps = float(ln[2])
        keep = (g, A, Wt, out, bias, rowbias, colscale, residual, qkv, ws, out16, ln_stats_out, ln)
        n_out = N // 2 if act == ACT_GEGLU else N
        a_bytes = (conv[0] * conv[1] * conv[2] * conv[3] * conv_stride * conv_stride if conv is not None else M * K) * 2
        if conv_up2:
            n_out = N // 4  # four phase blocks of Cout columns; 4 M output rows
        o_bytes = (4 * M if conv_up2 else M) * n_out * (4 if (qkv is None and out.dtype == torch.float32) else 2)
        desc = (f"{'conv' if conv is not None else 'lin'} M{M} N{N} K{K} "
                f"{'qkv' if qkv is not None else ('f32' if out.dtype == torch.float32 else 'f16')}"
                f"{' res' if residual is not None else ''}{' act%d' % act if act else ''}{' sk' if split_k != 1 else ''}"
                f"{' +f16' if out16 is not None else ''}{' hilo' if hilo else ''}{' s2' if conv_stride == 2 else ''}"
                f"{' +st' if ln_stats_out is not None else ''}{' ln' if ln is not None else ''}{' up2' if conv_up2 else ''}")
        sig = gemm_signature(conv is not None, M, N, K, "qkv" if qkv is not None else str(out.dtype).split(".")[-1], residual is not None, act,
                             "+st" if ln_stats_out is not None else ("+ln" if ln is not None else ("+up" if conv_up2 else "")))
        # algorithmic flops = those of the reference operation: the folded upsample convolution stands for conv3x3 over 4 M output pixels
        # (2 * 4M * Cout * 9C = 4.5 M N K with N = 4 Cout, K = 4 C) and executes 2 M N K
        meta = {"kernel": "gemm_tc_kernel", "flops": (4.5 if conv_up2 else 2.0) * M * N * K, "executed_flops": 2.0 * M * N * K * (3 if hilo else 1), "shape": (M, N, K, split_k), "desc": desc, "sig": sig, "can_split": ws is not None,
                "bytes": a_bytes + N * K * 2 + o_bytes + (M * n_out * 4 if residual is not None else 0) + (M * n_out * 2 if out16 is not None else 0)
                + (M * N // 4 if ln_stats_out is not None else 0) + (M * K // 4 if ln is not None else 0)}
        return self._bind("mvd_gemm_f16", (ctypes.byref(g),), keep, meta)

    def attn_self(self, q, k, vt, out, n_img, heads, seq, dhead, dpad, ldo, seq_valid=None):
        """seq_valid < seq: keys [seq_valid, seq) are masked (padded layout of a sequence that is not a multiple of 16)"""
        sv = seq if seq_valid is None else seq_valid
        return self._bind("mvd_attn_self_masked_f16", (_ptr(q, torch.float16), _ptr(k, torch.float16), _ptr(vt, torch.float16),
                                                       _ptr(out, torch.float16), n_img, heads, seq, sv, dhead, dpad, ldo),
                          (q, k, vt, out),
                          {"kernel": "attn_self_kernel", "flops": 4.0 * n_img * heads * seq * sv * dhead,
                           "desc": f"img{n_img} seq{seq} d{dhead}",
                           "bytes": 2.0 * n_img * heads * seq * (3 * dpad + dhead)})

    # ------------------------------------------------------------------ normalisation
    def groupnorm(self, x, gamma, beta, y, stats_ws, n_img, hw, C, eps, silu):
        return self._bind("mvd_groupnorm_f32_f16", (_ptr(x, torch.float32), _ptr(gamma, torch.float32),
                                                    _ptr(beta, torch.float32), _ptr(y, torch.float16),
                                                    _ptr(stats_ws, torch.float64) if stats_ws is not None else None, n_img, hw, C,
                                                    eps, int(silu)),
                          (x, gamma, beta, y, stats_ws), {"desc": f"img{n_img} hw{hw} C{C}", "bytes": 6.0 * n_img * hw * C})

    def groupnorm_hilo(self, x, gamma, beta, y, n_img, hw, C, eps, silu):
        """y fp16 [n_img*hw, 2C] = [hi | lo] (split-precision operand)"""
        return self._bind("mvd_groupnorm_hilo_f32_f16", (_ptr(x, torch.float32), _ptr(gamma, torch.float32), _ptr(beta, torch.float32),
                                                         _ptr(y, torch.float16), n_img, hw, C, eps, int(silu)),
                          (x, gamma, beta, y), {"kernel": "groupnorm_f32_f16", "desc": f"img{n_img} hw{hw} C{C} hilo", "bytes": 8.0 * n_img * hw * C})

    def groupnorm2(self, x1, C1, x2, C2, gamma, beta, y, n_img, hw, eps, silu):
        return self._bind("mvd_groupnorm2_f32_f16", (_ptr(x1, torch.float32), C1, _ptr(x2, torch.float32), C2, _ptr(gamma, torch.float32),
                                                     _ptr(beta, torch.float32), _ptr(y, torch.float16), n_img, hw, eps, int(silu)),
                          (x1, x2, gamma, beta, y), {"kernel": "groupnorm_f32_f16", "desc": f"img{n_img} hw{hw} C{C1}+{C2}",
                                                     "bytes": 6.0 * n_img * hw * (C1 + C2)})

    def layernorm(self, x, gamma, beta, y, rows, C, eps):
        return self._bind("mvd_layernorm_f32_f16", (_ptr(x, torch.float32), _ptr(gamma, torch.float32),
                                                    _ptr(beta, torch.float32), _ptr(y, torch.float16), rows, C, eps),
                          (x, gamma, beta, y), {"desc": f"rows{rows} C{C}", "bytes": 6.0 * rows * C})

    def layernorm_f32(self, x, gamma, beta, y, rows, C, eps, ldx=None, ldy=None):
        """fp32 -> fp32 LayerNorm over rows with pitches ldx / ldy (defaults C)"""
        return self._bind("mvd_layernorm_f32_f32", (_ptr(x, torch.float32), ldx if ldx is not None else C, _ptr(gamma, torch.float32),
                                                    _ptr(beta, torch.float32), _ptr(y, torch.float32), ldy if ldy is not None else C, rows, C, eps),
                          (x, gamma, beta, y), {"kernel": "layernorm_f32_f32", "desc": f"rows{rows} C{C}", "bytes": 8.0 * rows * C})

    def ln_modulate(self, x, shift, scale, y, rows, C, eps):
        return self._bind("mvd_ln_modulate_f32_f16", (_ptr(x, torch.float32), _ptr(shift, torch.float32),
                                                      _ptr(scale, torch.float32), _ptr(y, torch.float16), rows, C, eps),
                          (x, shift, scale, y))

    def softmax_rows(self, s, p, rows, cols, scale, ld_in=None, ld_out=None):
        return self._bind("mvd_softmax_rows_f32_f16", (_ptr(s, torch.float32), _ptr(p, torch.float16), rows, cols,
                                                       ld_in if ld_in is not None else cols, ld_out if ld_out is not None else cols,
                                                       float(scale)), (s, p), {"desc": f"rows{rows} cols{cols}", "bytes": 6.0 * rows * cols})

    # ------------------------------------------------------------------ data movement / elementwise
    def cast(self, x, y, n):
        return self._bind("mvd_cast_f32_f16", (_ptr(x, torch.float32), _ptr(y, torch.float16), n), (x, y),
                          {"desc": f"n{n}", "bytes": 6.0 * n})

    def concat(self, a, b, out, rows, C1, C2):
        return self._bind("mvd_concat_f32", (_ptr(a, torch.float32), _ptr(b, torch.float32), _ptr(out, torch.float32),
                                             rows, C1, C2), (a, b, out), {"desc": f"rows{rows} {C1}+{C2}", "bytes": 8.0 * rows * (C1 + C2)})

    def concat16(self, a, b, out, rows, C1, C2):
        return self._bind("mvd_concat_f32_f16", (_ptr(a, torch.float32), _ptr(b, torch.float32), _ptr(out, torch.float16),
                                                 rows, C1, C2), (a, b, out), {"desc": f"rows{rows} {C1}+{C2}", "bytes": 6.0 * rows * (C1 + C2)})

    def upsample2x(self, x, y, n_img, H, W, C):
        return self._bind("mvd_upsample2x_f32_f16", (_ptr(x, torch.float32), _ptr(y, torch.float16), n_img, H, W, C),
                          (x, y))

    def im2col_s2(self, x, y, n_img, H, W, C, pad_lo=1):
        return self._bind("mvd_im2col_s2_pad_f32_f16", (_ptr(x, torch.float32), _ptr(y, torch.float16), n_img, H, W, C, pad_lo),
                          (x, y))

    def gemv(self, x, W, bias, y, M, N, K, *, ldx=None, ldw=None, ldy=None, silu_in=False, silu_out=False):
        return self._bind("mvd_gemv_f16", (_ptr(x, torch.float32), ldx if ldx is not None else x.shape[-1],
                                           _ptr(W, torch.float16), ldw if ldw is not None else W.shape[-1],
                                           _ptr(bias, torch.float32), _ptr(y, torch.float32),
                                           ldy if ldy is not None else y.shape[-1], M, N, K, int(silu_in),
                                           int(silu_out)), (x, W, bias, y))

    def gemv_grouped(self, x, K, jobs, *, silu_in=False):
        """jobs: list of (W fp16 [N, ldw], bias fp32 [N] or None, y fp32 [N]); all read the single input row x[:K]."""
        rows, col0 = [], 0
        for W, bias, y in jobs:
            N = y.numel()
            rows.append([_ptr(W, torch.float16), _ptr(bias, torch.float32) or 0, _ptr(y, torch.float32), N | (W.shape[-1] << 32), col0])
            col0 += N
        table = torch.tensor(rows, dtype=torch.int64).to(self.device)
        return self._bind("mvd_gemv_grouped_f16", (_ptr(x, torch.float32), K, int(silu_in), _ptr(table, torch.int64), len(jobs), col0),
                          (x, table, jobs), {"desc": f"jobs{len(jobs)} cols{col0} K{K}", "bytes": 2.0 * col0 * K})

    def timestep_embedding(self, t_dev, freqs, out, dim):
        return self._bind("mvd_timestep_embedding", (_ptr(t_dev, torch.float32), _ptr(freqs, torch.float32),
                                                     _ptr(out, torch.float32), dim), (t_dev, freqs, out))

    def unet_input(self, noisy, cond, cond_batched, cond_scale, out, n_views, n_img, hw, Cpad, hilo=False):
        return self._bind("mvd_unet_input_f16", (_ptr(noisy, torch.float32), _ptr(cond, torch.float32),
                                                 int(cond_batched), _ptr(cond_scale, torch.float32),
                                                 _ptr(out, torch.float16), n_views, n_img, hw, Cpad, int(hilo)),
                          (noisy, cond, cond_scale, out))

    def cfg_ddim(self, head, ld, two_branch, coef, xt, noise, eps_out, x_prev, x0_out, n_views, hw):
        return self._bind("mvd_cfg_ddim", (_ptr(head, torch.float32), ld, int(two_branch), _ptr(coef, torch.float32),
                                           _ptr(xt, torch.float32), _ptr(noise, torch.float32),
                                           _ptr(eps_out, torch.float32), _ptr(x_prev, torch.float32),
                                           _ptr(x0_out, torch.float32), n_views, hw),
                          (head, coef, xt, noise, eps_out, x_prev, x0_out))

    def nchw_to_rows(self, x, y, n_img, C, hw):
        return self._bind("mvd_nchw_to_rows_f32", (_ptr(x, torch.float32), _ptr(y, torch.float32), n_img, C, hw), (x, y))

    def rows_to_nchw(self, x, y, n_img, C, ld, hw):
        return self._bind("mvd_rows_to_nchw_f32", (_ptr(x, torch.float32), _ptr(y, torch.float32), n_img, C, ld, hw),
                          (x, y))

    def nchw_to_nhwc16(self, x, y, n_img, C, hw, Cpad):
        return self._bind("mvd_nchw_to_nhwc_f16", (_ptr(x, torch.float32), _ptr(y, torch.float16), n_img, C, hw, Cpad), (x, y))

    def gather_rows(self, table, row_len, idx_dev, out):
        return self._bind("mvd_gather_rows_f32", (_ptr(table, torch.float32), row_len, _ptr(idx_dev, torch.int32),
                                                  _ptr(out, torch.float32)), (table, idx_dev, out))

    def increment(self, counter, delta):
        return self._bind("mvd_increment_i32", (_ptr(counter, torch.int32), delta), (counter,))

    # ------------------------------------------------------------------ GridAttn
    def gridattn_prep(self, noisy, input_latent, depth_override, depth_eps, scal, Wz, bz, feat, zdepth, n_views, S, D,
                      depth_scale, depth_shift):
        return self._bind("mvd_gridattn_prep", (_ptr(noisy, torch.float32), _ptr(input_latent, torch.float32),
                                                _ptr(depth_override, torch.float32), _ptr(depth_eps, torch.float32),
                                                _ptr(scal, torch.float32), _ptr(Wz, torch.float32),
                                                _ptr(bz, torch.float32), _ptr(feat, torch.float16),
                                                _ptr(zdepth, torch.float32), n_views, S, D, depth_scale, depth_shift),
                          (noisy, input_latent, depth_override, depth_eps, scal, Wz, bz, feat, zdepth))

    def gridattn_tokens(self, feat, zdepth, cams, mask, freqs, ndc_grid, tokens, n_views, S, D, q_first, q_count):
        return self._bind("mvd_gridattn_tokens", (_ptr(feat, torch.float16), _ptr(zdepth, torch.float32),
                                                  _ptr(cams, torch.float32), _ptr(mask, torch.float32),
                                                  _ptr(freqs, torch.float32), _ptr(ndc_grid, torch.float32),
                                                  _ptr(tokens, torch.float16), n_views, S, D, q_first, q_count),
                          (feat, zdepth, cams, mask, freqs, ndc_grid, tokens))

    def view_attention(self, qkv, out, P, V, heads, hd):
        return self._bind("mvd_view_attention_f16", (_ptr(qkv, torch.float16), _ptr(out, torch.float16), P, V, heads, hd),
                          (qkv, out))

    def dit_fold_gates(self, jobs):
        """jobs: list of (W fp16 [N, K], gate fp32 [N], bias fp32 [N], W_out fp16 [N, K], b_out fp32 [N]): W_out = gate[:, None] * W, b_out = gate * b"""
        arr = (_lib.FoldJob * len(jobs))()
        for j, (W, gate, bias, Wo, bo) in zip(arr, jobs):
            j.w, j.gate, j.bias, j.w_out, j.b_out = (_ptr(W, torch.float16), _ptr(gate, torch.float32), _ptr(bias, torch.float32),
                                                     _ptr(Wo, torch.float16), _ptr(bo, torch.float32))
            j.N, j.K = W.shape[0], W.shape[1]
        return self._bind("mvd_dit_fold_gates", (arr, len(jobs)), (arr, jobs))

    def gridattn_dit(self, tokens, token_k, w_pre, b_pre, layers, pool_w, pool_b, pooled, R, V, eps, x_out=None):
        """The aggregation transformer as one kernel (include/mvd_b200.h, mvd_gridattn_dit_f16).  layers: list of dicts with the
        mvd_dit_layer field names (w_qkv in head order, w_proj / w_fc2 / b_proj / b_fc2 gate-folded)."""
        a = _lib.DitArgs()
        a.R, a.V, a.layers, a.token_k, a.token_ld = R, V, len(layers), token_k, tokens.shape[-1]
        a.tokens, a.w_pre, a.w_pre_ld, a.b_pre = _ptr(tokens, torch.float16), _ptr(w_pre, torch.float16), w_pre.shape[-1], _ptr(b_pre, torch.float32)
        for i, lay in enumerate(layers):
            for name, _ in _lib.DitLayer._fields_:
                setattr(a.layer[i], name, _ptr(lay[name], torch.float16 if name.startswith("w_") else torch.float32))
        a.pool_w, a.pool_b, a.pooled, a.x_out, a.eps = (_ptr(pool_w, torch.float32), _ptr(pool_b, torch.float32), _ptr(pooled, torch.float16),
                                                         _ptr(x_out, torch.float32), eps)
        nl = len(layers)
        flops = 2.0 * R * (token_k * 256 + nl * (768 * 256 + 256 * 256 + 2 * 512 * 256))
        return self._bind("mvd_gridattn_dit_f16", (ctypes.byref(a),), (a, tokens, w_pre, b_pre, layers, pool_w, pool_b, pooled, x_out),
                          {"kernel": "gridattn_dit", "desc": f"rows{R} V{V} layers{nl}", "flops": flops, "bytes": 2.0 * R * tokens.shape[-1]})

    def view_pool(self, x, w, b, out, P, V, C):
        return self._bind("mvd_view_pool_f16", (_ptr(x, torch.float32), _ptr(w, torch.float32), _ptr(b, torch.float32),
                                                _ptr(out, torch.float16), P, V, C), (x, w, b, out))

    def frustum_pool(self, inp, out, n_img, S, D, C, factor):
        return self._bind("mvd_frustum_pool_f16", (_ptr(inp, torch.float16), _ptr(out, torch.float16), n_img, S, D, C,
                                                   factor), (inp, out))

    def pixel_cross_attn(self, q, kv, out, M, D, heads, dhead):
        return self._bind("mvd_pixel_cross_attn_f16", (_ptr(q, torch.float16), _ptr(kv, torch.float16),
                                                       _ptr(out, torch.float16), M, D, heads, dhead), (q, kv, out))

    # ------------------------------------------------------------------ training (ABI 15): fp32 passes with saved statistics + backward
    def layernorm_fwd(self, x, gamma, beta, y, stats, rows, C, eps):
        """y fp32 [rows, C], stats fp32 [rows, 2] = (mean, rstd); gamma = beta = None: no affine part"""
        return self._bind("mvd_layernorm_fwd_f32", (_ptr(x, torch.float32), _ptr(gamma, torch.float32), _ptr(beta, torch.float32),
                                                    _ptr(y, torch.float32), _ptr(stats, torch.float32), rows, C, eps),
                          (x, gamma, beta, y, stats), {"kernel": "train_layernorm", "desc": f"rows{rows} C{C}", "bytes": 8.0 * rows * C})

    def layernorm_bwd(self, dy, x, gamma, stats, dx, dgamma, dbeta, rows, C):
        return self._bind("mvd_layernorm_bwd_f32", (_ptr(dy, torch.float32), _ptr(x, torch.float32), _ptr(gamma, torch.float32),
                                                    _ptr(stats, torch.float32), _ptr(dx, torch.float32), _ptr(dgamma, torch.float32),
                                                    _ptr(dbeta, torch.float32), rows, C),
                          (dy, x, gamma, stats, dx, dgamma, dbeta), {"kernel": "train_layernorm", "desc": f"bwd rows{rows} C{C}",
                                                                     "bytes": (12.0 + (8.0 if dgamma is not None else 0.0)) * rows * C})

    def groupnorm_fwd(self, x, gamma, beta, y, stats, ws, n_img, hw, C, eps, silu):
        """x, y fp32 [n_img, hw, C]; stats fp32 [n_img, 32, 2]; ws: uint8 scratch of n_img * GN_WS_BYTES_PER_IMAGE bytes"""
        if ws.numel() < n_img * GN_WS_BYTES_PER_IMAGE:
            raise MvdError("groupnorm_fwd: workspace too small")
        return self._bind("mvd_groupnorm_fwd_f32", (_ptr(x, torch.float32), _ptr(gamma, torch.float32), _ptr(beta, torch.float32),
                                                    _ptr(y, torch.float32), _ptr(stats, torch.float32), _ptr(ws, torch.uint8), n_img, hw, C,
                                                    eps, int(silu)),
                          (x, gamma, beta, y, stats, ws), {"kernel": "train_groupnorm", "desc": f"img{n_img} hw{hw} C{C}", "bytes": 12.0 * n_img * hw * C})

    def groupnorm_bwd(self, dy, x, gamma, beta, stats, dx, dgamma, dbeta, ws, n_img, hw, C, silu):
        if ws.numel() < n_img * GN_WS_BYTES_PER_IMAGE:
            raise MvdError("groupnorm_bwd: workspace too small")
        return self._bind("mvd_groupnorm_bwd_f32", (_ptr(dy, torch.float32), _ptr(x, torch.float32), _ptr(gamma, torch.float32),
                                                    _ptr(beta, torch.float32), _ptr(stats, torch.float32), _ptr(dx, torch.float32),
                                                    _ptr(dgamma, torch.float32), _ptr(dbeta, torch.float32), _ptr(ws, torch.uint8),
                                                    n_img, hw, C, int(silu)),
                          (dy, x, gamma, beta, stats, dx, dgamma, dbeta, ws),
                          {"kernel": "train_groupnorm", "desc": f"bwd img{n_img} hw{hw} C{C}", "bytes": 20.0 * n_img * hw * C})

    def act_fwd(self, x, y, rows, cols, mode):
        """mode ACT_GELU / ACT_SILU: elementwise over rows * cols; ACT_GEGLU: x [rows, 2 cols] -> y [rows, cols]"""
        return self._bind("mvd_act_fwd_f32", (_ptr(x, torch.float32), _ptr(y, torch.float32), rows, cols, mode), (x, y),
                          {"kernel": "train_act", "desc": f"rows{rows} cols{cols} mode{mode}", "bytes": (12.0 if mode == ACT_GEGLU else 8.0) * rows * cols})

    def act_bwd(self, dy, x, dx, rows, cols, mode):
        return self._bind("mvd_act_bwd_f32", (_ptr(dy, torch.float32), _ptr(x, torch.float32), _ptr(dx, torch.float32), rows, cols, mode),
                          (dy, x, dx), {"kernel": "train_act", "desc": f"bwd rows{rows} cols{cols} mode{mode}",
                                        "bytes": (20.0 if mode == ACT_GEGLU else 12.0) * rows * cols})

    def bilinear_gather_fwd(self, fmap, xy, out, V, H, W, C, P):
        """fmap fp32 [V, H, W, C] (channels-last), xy fp32 [V, P, 2] grid_sample coordinates, out fp32 [V, P, C] (ABI 16)"""
        return self._bind("mvd_bilinear_gather_fwd_f32", (_ptr(fmap, torch.float32), _ptr(xy, torch.float32), _ptr(out, torch.float32), V, H, W, C, P),
                          (fmap, xy, out), {"kernel": "train_gather", "desc": f"V{V} {H}x{W}x{C} P{P}", "bytes": 4.0 * V * P * C})

    def bilinear_gather_bwd(self, dout, xy, dfmap, V, H, W, C, P):
        return self._bind("mvd_bilinear_gather_bwd_f32", (_ptr(dout, torch.float32), _ptr(xy, torch.float32), _ptr(dfmap, torch.float32), V, H, W, C, P),
                          (dout, xy, dfmap), {"kernel": "train_gather", "desc": f"bwd V{V} {H}x{W}x{C} P{P}", "bytes": 4.0 * V * P * C})
