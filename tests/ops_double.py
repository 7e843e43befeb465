"""TEST DOUBLE for mvdfusion_b200.ops.NativeOps — test infrastructure, never imported by the product.

It emulates, with plain torch on CPU tensors, the documented semantics of every C-ABI entry point
(include/mvd_b200.h), including the fp16 rounding of fp16 outputs.  Purpose: on a box without a GPU, run the
product's host-side logic (engine emitters, weight packing, buffer arena, program order, step tables, view sharding)
against the oracle, so that the only thing left to verify on the B200 is the kernels themselves.
"""
import sys

import torch
import torch.nn.functional as F

from mvdfusion_b200.ops import ACT_GEGLU, ACT_GELU, ACT_NONE, ACT_SILU


def geglu_permutation(inner, tile_n):
    """mirror of mvd_geglu_row_permutation (csrc/gemm.cu); tests/test_lib_host.py checks it against the library"""
    half = tile_n // 2
    perm = []
    for r in range(2 * inner):
        tile, w = divmod(r, tile_n)
        perm.append(tile * half + w if w < half else inner + tile * half + (w - half))
    return torch.tensor(perm, dtype=torch.long)


class TorchOpsDouble:
    def __init__(self, device="cpu"):
        self.device = torch.device(device)
        self.calls = 0

    def empty(self, shape, dtype):
        # poison (NaN for floats, 0xFF bytes for the arena's raw buffers): catches reads of unwritten memory
        return torch.full(shape, float("nan") if dtype.is_floating_point else 255, dtype=dtype)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def geglu_permutation(self, inner, tile_n):
        return geglu_permutation(inner, tile_n)

    def _call(self, fn):
        def run(stream):
            self.calls += 1
            fn()
        run.name = sys._getframe(1).f_code.co_name  # the emulated op (program-structure tests count launches by kind)
        return run

    # ------------------------------------------------------------------ GEMM / conv
    def gemm(self, A, Wt, out, M, N, K, *, lda=None, ldw=None, ldc=None, bias=None, rowbias=None, rows_per_group=1,
             colscale=None, residual=None, ldr=0, act=ACT_NONE, conv=None, qkv=None, split_k=1, tile_n=0, ws=None, cta_pair=0,
             out16=None, ld16=None, hilo=False, out16_lo=0, a_lo_off=0, conv_stride=1, conv_no_pad_lo=False,
             ln_stats_out=None, ln=None, conv_up2=False):
        assert A.dtype == torch.float16 and Wt.dtype == torch.float16
        if conv_up2:
            return self._gemm_up2(A, Wt, out, M, N, K, lda=lda, ldw=ldw, ldc=ldc, bias=bias, conv=conv, out16=out16, ld16=ld16, out16_lo=out16_lo,
                                  plain=(rowbias is None and colscale is None and residual is None and act == ACT_NONE and qkv is None
                                         and not hilo and conv_stride == 1 and not conv_no_pad_lo and ln is None and ln_stats_out is None),
                                  tile_n=tile_n)
        # ABI 13 (LayerNorm between two GEMMs): the producer leaves per-chunk (sum, sum of squares), the consumer folds the normalisation
        assert ln_stats_out is None or (out.dtype == torch.float32 and qkv is None and act != ACT_GEGLU and N % 32 == 0 and split_k in (0, 1)
                                        and K <= 1536 and ln_stats_out.numel() >= (N // 32) * M * 2)
        assert ln is None or (conv is None and not hilo and K % 32 == 0 and N % 32 == 0 and (qkv is not None or act == ACT_GEGLU)
                              and split_k in (0, 1) and ln[0].numel() >= (K // 32) * M * 2)
        assert out16 is None or (out.dtype == torch.float32 and qkv is None and act != ACT_GEGLU)
        ldw_ = ldw if ldw is not None else Wt.shape[-1]

        def fn():
            Wfull = Wt.reshape(-1)[: N * ldw_].reshape(N, ldw_)
            W = Wfull[:, :K].float()
            if conv is not None:
                n_img, H, Wd, C = conv
                assert K == 9 * C and M == n_img * H * Wd
                Ct = 2 * C if hilo else C
                st = conv_stride
                Hi, Wi = H * st, Wd * st           # H, Wd are the OUTPUT extent; the image holds (st H) x (st Wd) pixels
                pitch = lda if lda else Ct
                lo_pad, hi_pad = (0 if conv_no_pad_lo else 1), 2

                def im2col(x):
                    xp = F.pad(x, (0, 0, lo_pad, hi_pad, lo_pad, hi_pad))
                    return torch.cat([xp[:, ky:ky + Hi:st, kx:kx + Wi:st, :][:, :H, :Wd] for ky in range(3) for kx in range(3)], dim=-1).reshape(M, K)
                xall = torch.as_strided(A, (n_img, Hi, Wi, Ct), (Hi * Wi * pitch, Wi * pitch, pitch, 1), A.storage_offset()).float()
                a = im2col(xall[..., :C])
                a_lo = im2col(xall[..., C:]) if hilo else None
            else:
                lda_ = lda if lda is not None else A.shape[-1]
                a = torch.as_strided(A, (M, K), (lda_, 1), A.storage_offset()).float()  # A may be a column window of a wider buffer
                a_lo = torch.as_strided(A, (M, K), (lda_, 1), A.storage_offset() + (a_lo_off or K)).float() if hilo else None
            acc = a @ W.t()
            if hilo:  # A_hi W_hi + A_lo W_hi + A_hi W_lo
                assert K % 64 == 0
                acc = acc + a_lo @ W.t() + a @ Wfull[:, K:2 * K].float().t()
            if ln is not None:  # statistics from the producer's chunk sums, E[x^2] - mean^2 form, as the kernel does
                st = ln[0].reshape(-1)[: (K // 32) * M * 2].reshape(K // 32, M, 2).sum(dim=0)
                mu = st[:, 0:1] / K
                var = (st[:, 1:2] / K - mu * mu).clamp_min(0.0)
                acc = torch.rsqrt(var + ln[2]) * (acc - mu * ln[1].reshape(-1)[:N].float())
            if bias is not None:
                acc = acc + bias.reshape(-1)[:N]
            if rowbias is not None:
                groups = (M + rows_per_group - 1) // rows_per_group
                rb = rowbias.reshape(-1)[: groups * N].reshape(groups, N)
                acc = acc + rb.repeat_interleave(rows_per_group, dim=0)[:M]
            if act == ACT_GELU:
                acc = F.gelu(acc)
            elif act == ACT_SILU:
                acc = F.silu(acc)
            elif act == ACT_GEGLU:
                bn = tile_n if tile_n else 256
                t = acc.reshape(M, N // bn, bn)
                acc = (t[..., : bn // 2] * F.gelu(t[..., bn // 2:])).reshape(M, N // 2)
            if colscale is not None:
                assert act != ACT_GEGLU
                acc = acc * colscale.reshape(-1)[:N]
            if residual is not None:
                acc = acc + residual.reshape(-1)[: M * ldr].reshape(M, ldr)[:, : acc.shape[1]]
            if qkv is not None:
                heads, d, dpad, seq = qkv["heads"], qkv["dhead"], qkv["dpad"], qkv["seq"]
                n_img = M // seq
                t = acc.reshape(n_img, seq, 3, heads, d).permute(2, 0, 3, 1, 4)  # [3, img, h, seq, d]
                q = out.reshape(-1)[: n_img * heads * seq * dpad].reshape(n_img, heads, seq, dpad)
                k = qkv["out_k"].reshape(-1)[: n_img * heads * seq * dpad].reshape(n_img, heads, seq, dpad)
                vt = qkv["out_vt"].reshape(-1)[: n_img * heads * seq * dpad].reshape(n_img, heads, dpad, seq)
                q[..., :d] = t[0].half()
                k[..., :d] = t[1].half()
                vt[:, :, :d, :] = t[2].transpose(-1, -2).half()
                return
            ldc_ = ldc if ldc is not None else out.shape[-1]
            o = out.reshape(-1)[: M * ldc_].reshape(M, ldc_)
            o[:, : acc.shape[1]] = acc.to(out.dtype)
            if ln_stats_out is not None:
                ch = acc.reshape(M, N // 32, 32)
                ln_stats_out.reshape(-1)[: (N // 32) * M * 2].reshape(N // 32, M, 2).copy_(
                    torch.stack([ch.sum(-1), (ch * ch).sum(-1)], dim=-1).permute(1, 0, 2))
            if out16 is not None:
                ld16_ = ld16 if ld16 is not None else out16.shape[-1]
                # out16 may be a column window of a wider buffer: address it from its first element with the row pitch
                o16 = torch.as_strided(out16, (M, N), (ld16_, 1), out16.storage_offset()) if out16.dim() else out16
                o16.copy_(acc.half())
                if out16_lo:
                    lo = torch.as_strided(out16, (M, N), (ld16_, 1), out16.storage_offset() + out16_lo)
                    lo.copy_((acc - acc.half().float()).half())
        return self._call(fn)

    def _gemm_up2(self, A, Wt, out, M, N, K, *, lda, ldw, ldc, bias, conv, out16, ld16, out16_lo, plain, tile_n):
        """ABI 14: nearest x2 upsample folded into the convolution — four 2 x 2 phase convolutions of the SOURCE image."""
        n_img, H, Wd, C = conv
        Co = N // 4
        assert plain and conv is not None and N % 4 == 0 and K == 4 * C and M == n_img * H * Wd and H & (H - 1) == 0 and Wd & (Wd - 1) == 0
        assert tile_n == 0 or (tile_n % 32 == 0 and tile_n <= 256 and Co % tile_n == 0)
        ldw_ = ldw if ldw is not None else Wt.shape[-1]

        def fn():
            W = Wt.reshape(-1)[: N * ldw_].reshape(N, ldw_)[:, :K].float()
            pitch = lda if lda else C
            x = torch.as_strided(A, (n_img, H, Wd, C), (H * Wd * pitch, Wd * pitch, pitch, 1), A.storage_offset()).float()
            xp = F.pad(x, (0, 0, 1, 1, 1, 1))  # source pixel (y, x) at [y + 1, x + 1]
            ldc_ = ldc if ldc is not None else out.shape[-1]
            o = torch.as_strided(out, (n_img, 2 * H, 2 * Wd, Co), (4 * H * Wd * ldc_, 2 * Wd * ldc_, ldc_, 1), out.storage_offset())
            res = torch.empty(n_img, 2 * H, 2 * Wd, Co)
            for py in range(2):
                for px in range(2):
                    ph = 2 * py + px
                    # taps (a, b): source pixel (y + py - 1 + a, x + px - 1 + b)
                    cols = torch.cat([xp[:, py + a:py + a + H, px + b_:px + b_ + Wd, :] for a in range(2) for b_ in range(2)], dim=-1).reshape(M, K)
                    acc = cols @ W[ph * Co:(ph + 1) * Co].t()
                    if bias is not None:
                        acc = acc + bias.reshape(-1)[ph * Co:(ph + 1) * Co]
                    res[:, py::2, px::2, :] = acc.reshape(n_img, H, Wd, Co)
            o.copy_(res.to(out.dtype))
            if out16 is not None:
                ld16_ = ld16 if ld16 is not None else out16.shape[-1]
                o16 = torch.as_strided(out16, (n_img, 2 * H, 2 * Wd, Co), (4 * H * Wd * ld16_, 2 * Wd * ld16_, ld16_, 1), out16.storage_offset())
                o16.copy_(res.half())
                if out16_lo:
                    lo = torch.as_strided(out16, (n_img, 2 * H, 2 * Wd, Co), (4 * H * Wd * ld16_, 2 * Wd * ld16_, ld16_, 1), out16.storage_offset() + out16_lo)
                    lo.copy_((res - res.half().float()).half())
        return self._call(fn)

    def attn_self(self, q, k, vt, out, n_img, heads, seq, dhead, dpad, ldo, seq_valid=None):
        def fn():
            n = n_img * heads * seq * dpad
            Q = q.reshape(-1)[:n].reshape(n_img, heads, seq, dpad).float()
            Kk = k.reshape(-1)[:n].reshape(n_img, heads, seq, dpad).float()
            V = vt.reshape(-1)[:n].reshape(n_img, heads, dpad, seq).float().transpose(-1, -2)
            kd = (dhead + 15) // 16 * 16
            s = (Q[..., :kd] @ Kk[..., :kd].transpose(-1, -2)) * dhead ** -0.5
            if seq_valid is not None and seq_valid < seq:
                s[..., seq_valid:] = float("-inf")
            p = s.softmax(-1)
            o = (p @ V)[..., :dhead]
            out.reshape(-1)[: n_img * seq * ldo].reshape(n_img, seq, ldo)[:, :, : heads * dhead] = \
                o.permute(0, 2, 1, 3).reshape(n_img, seq, heads * dhead).half()
        return self._call(fn)

    # ------------------------------------------------------------------ normalisation
    def groupnorm(self, x, gamma, beta, y, stats_ws, n_img, hw, C, eps, silu):
        def fn():
            v = x.reshape(-1)[: n_img * hw * C].reshape(n_img, hw, C).permute(0, 2, 1)
            o = F.group_norm(v, 32, gamma, beta, eps)
            if silu:
                o = F.silu(o)
            y.reshape(-1)[: n_img * hw * C].copy_(o.permute(0, 2, 1).reshape(-1).half())
        return self._call(fn)

    def groupnorm_hilo(self, x, gamma, beta, y, n_img, hw, C, eps, silu):
        def fn():
            v = x.reshape(-1)[: n_img * hw * C].reshape(n_img, hw, C).permute(0, 2, 1)
            o = F.group_norm(v, 32, gamma, beta, eps)
            if silu:
                o = F.silu(o)
            o = o.permute(0, 2, 1).reshape(n_img * hw, C)
            yy = y.reshape(-1)[: n_img * hw * 2 * C].reshape(n_img * hw, 2 * C)
            yy[:, :C] = o.half()
            yy[:, C:] = (o - o.half().float()).half()
        return self._call(fn)

    def groupnorm2(self, x1, C1, x2, C2, gamma, beta, y, n_img, hw, eps, silu):
        def fn():
            a = x1.reshape(-1)[: n_img * hw * C1].reshape(n_img, hw, C1)
            b = x2.reshape(-1)[: n_img * hw * C2].reshape(n_img, hw, C2)
            v = torch.cat([a, b], dim=2).permute(0, 2, 1)
            o = F.group_norm(v, 32, gamma, beta, eps)
            if silu:
                o = F.silu(o)
            y.reshape(-1)[: n_img * hw * (C1 + C2)].copy_(o.permute(0, 2, 1).reshape(-1).half())
        return self._call(fn)

    def layernorm(self, x, gamma, beta, y, rows, C, eps):
        def fn():
            v = x.reshape(-1)[: rows * C].reshape(rows, C)
            y.reshape(-1)[: rows * C].copy_(F.layer_norm(v, (C,), gamma, beta, eps).reshape(-1).half())
        return self._call(fn)

    def layernorm_f32(self, x, gamma, beta, y, rows, C, eps, ldx=None, ldy=None):
        lx, ly = ldx if ldx is not None else C, ldy if ldy is not None else C

        def fn():
            v = torch.as_strided(x, (rows, C), (lx, 1), x.storage_offset()).clone()
            torch.as_strided(y, (rows, C), (ly, 1), y.storage_offset()).copy_(F.layer_norm(v, (C,), gamma, beta, eps))
        return self._call(fn)

    def softmax_rows(self, s, p, rows, cols, scale, ld_in=None, ld_out=None):
        li, lo = ld_in if ld_in is not None else cols, ld_out if ld_out is not None else cols

        def fn():
            x = s.reshape(-1)[: rows * li].reshape(rows, li)[:, :cols].float()
            p.reshape(-1)[: rows * lo].reshape(rows, lo)[:, :cols] = torch.softmax(x * scale, dim=-1).half()
        return self._call(fn)

    def ln_modulate(self, x, shift, scale, y, rows, C, eps):
        def fn():
            v = x.reshape(-1)[: rows * C].reshape(rows, C)
            o = F.layer_norm(v, (C,), None, None, eps) * (1 + scale.reshape(-1)[:C]) + shift.reshape(-1)[:C]
            y.reshape(-1)[: rows * C].copy_(o.reshape(-1).half())
        return self._call(fn)

    # ------------------------------------------------------------------ data movement
    def cast(self, x, y, n):
        return self._call(lambda: y.reshape(-1)[:n].copy_(x.reshape(-1)[:n].half()))

    def concat(self, a, b, out, rows, C1, C2):
        def fn():
            o = out.reshape(-1)[: rows * (C1 + C2)].reshape(rows, C1 + C2)
            o[:, :C1] = a.reshape(-1)[: rows * C1].reshape(rows, C1)
            o[:, C1:] = b.reshape(-1)[: rows * C2].reshape(rows, C2)
        return self._call(fn)

    def concat16(self, a, b, out, rows, C1, C2):
        def fn():
            o = out.reshape(-1)[: rows * (C1 + C2)].reshape(rows, C1 + C2)
            o[:, :C1] = a.reshape(-1)[: rows * C1].reshape(rows, C1).half()
            o[:, C1:] = b.reshape(-1)[: rows * C2].reshape(rows, C2).half()
        return self._call(fn)

    def upsample2x(self, x, y, n_img, H, W, C):
        def fn():
            v = x.reshape(-1)[: n_img * H * W * C].reshape(n_img, H, W, C)
            o = v.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
            y.reshape(-1)[: o.numel()].copy_(o.reshape(-1).half())
        return self._call(fn)

    def im2col_s2(self, x, y, n_img, H, W, C, pad_lo=1):
        def fn():
            v = x.reshape(-1)[: n_img * H * W * C].reshape(n_img, H, W, C)
            vp = F.pad(v, (0, 0, pad_lo, 2 - pad_lo, pad_lo, 2 - pad_lo))
            cols = [vp[:, ky:ky + H:2, kx:kx + W:2, :] for ky in range(3) for kx in range(3)]
            o = torch.cat(cols, dim=-1)
            y.reshape(-1)[: o.numel()].copy_(o.reshape(-1).half())
        return self._call(fn)

    def gemv(self, x, W, bias, y, M, N, K, *, ldx=None, ldw=None, ldy=None, silu_in=False, silu_out=False):
        ldx_ = ldx if ldx is not None else x.shape[-1]
        ldw_ = ldw if ldw is not None else W.shape[-1]
        ldy_ = ldy if ldy is not None else y.shape[-1]

        def fn():
            a = x.reshape(-1)[: M * ldx_].reshape(M, ldx_)[:, :K]
            if silu_in:
                a = F.silu(a)
            o = a @ W.reshape(-1)[: N * ldw_].reshape(N, ldw_)[:, :K].float().t()
            if bias is not None:
                o = o + bias.reshape(-1)[:N]
            if silu_out:
                o = F.silu(o)
            y.reshape(-1)[: M * ldy_].reshape(M, ldy_)[:, :N] = o
        return self._call(fn)

    def gemv_grouped(self, x, K, jobs, *, silu_in=False):
        def fn():
            a = x.reshape(-1)[:K]
            if silu_in:
                a = F.silu(a)
            for W, bias, y in jobs:
                o = W[:, :K].float() @ a
                if bias is not None:
                    o = o + bias.reshape(-1)[: o.shape[0]]
                y.reshape(-1)[: o.shape[0]].copy_(o)
        return self._call(fn)

    def timestep_embedding(self, t_dev, freqs, out, dim):
        def fn():
            a = t_dev.reshape(-1)[0] * freqs
            out.reshape(-1)[:dim].copy_(torch.cat([torch.cos(a), torch.sin(a)])[:dim])
        return self._call(fn)

    def unet_input(self, noisy, cond, cond_batched, cond_scale, out, n_views, n_img, hw, Cpad, hilo=False):
        def fn():
            o = out.reshape(-1)[: n_img * hw * Cpad].reshape(n_img, hw, Cpad)
            o.zero_()
            nz = noisy.reshape(-1)[: n_views * 5 * hw].reshape(n_views, 5, hw)
            for img in range(n_img):
                view = img % n_views
                v = torch.zeros(hw, 10)
                v[:, :5] = nz[view].t()
                if img < n_views:
                    c = cond.reshape(-1, 5, hw)[view if cond_batched else 0].clone()
                    if cond_scale is not None:
                        c = c * cond_scale[view]
                    c[:4] = c[:4] / 0.18215
                    v[:, 5:10] = c.t()
                o[img, :, :10] = v.half()
                if hilo:  # [hi | lo | hi]: the three K segments of the split-precision stem conv as plain channels
                    o[img, :, 10:20] = (v - v.half().float()).half()
                    o[img, :, 20:30] = v.half()
        return self._call(fn)

    def cfg_ddim(self, head, ld, two_branch, coef, xt, noise, eps_out, x_prev, x0_out, n_views, hw):
        def fn():
            n_img = n_views * (2 if two_branch else 1)
            h = head.reshape(-1)[: n_img * hw * ld].reshape(n_img, hw, ld)[:, :, :5].permute(0, 2, 1)
            e = h[:n_views]
            if two_branch:
                e = h[n_views:] + coef[5] * (h[:n_views] - h[n_views:])
            if eps_out is not None:
                eps_out.reshape(-1)[: e.numel()].copy_(e.reshape(-1))
            if xt is not None:
                a_t, a_prev, somat, sigma = coef[0], coef[1], coef[2], coef[3]
                x = xt.reshape(n_views, 5, hw).clone()
                x0 = (x - somat * e) / a_t.sqrt()
                xp = a_prev.sqrt() * x0 + torch.clamp(1.0 - a_prev - sigma * sigma, min=1e-7).sqrt() * e
                if float(coef[4]) != 0.0:
                    xp = xp + sigma * noise.reshape(n_views, 5, hw)
                x_prev.reshape(n_views, 5, hw).copy_(xp)
                if x0_out is not None:
                    x0_out.reshape(n_views, 5, hw).copy_(x0)
        return self._call(fn)

    def nchw_to_rows(self, x, y, n_img, C, hw):
        return self._call(lambda: y.reshape(-1)[: n_img * C * hw].copy_(x.reshape(n_img, C, hw).permute(0, 2, 1).reshape(-1)))

    def rows_to_nchw(self, x, y, n_img, C, ld, hw):
        return self._call(lambda: y.reshape(-1)[: n_img * C * hw].copy_(
            x.reshape(-1)[: n_img * hw * ld].reshape(n_img, hw, ld)[:, :, :C].permute(0, 2, 1).reshape(-1)))

    def nchw_to_nhwc16(self, x, y, n_img, C, hw, Cpad):
        def fn():
            o = y.reshape(-1)[: n_img * hw * Cpad].reshape(n_img, hw, Cpad)
            o.zero_()
            o[:, :, :C] = x.reshape(n_img, C, hw).permute(0, 2, 1).half()
        return self._call(fn)

    def gather_rows(self, table, row_len, idx_dev, out):
        return self._call(lambda: out.reshape(-1)[:row_len].copy_(table.reshape(-1, row_len)[int(idx_dev[0])]))

    def increment(self, counter, delta):
        return self._call(lambda: counter.add_(delta))

    # ------------------------------------------------------------------ GridAttn
    def gridattn_prep(self, noisy, input_latent, depth_override, depth_eps, scal, Wz, bz, feat, zdepth, n_views, S, D,
                      depth_scale, depth_shift):
        def fn():
            hw = S * S
            lat = torch.cat([noisy.reshape(n_views, 5, hw), input_latent.reshape(-1, 5, hw)[:1]])
            f = F.gelu(lat.permute(0, 2, 1) @ Wz.reshape(256, 5).t() + bz)
            feat.reshape(-1)[: f.numel()].copy_(f.reshape(-1).half())
            mean = depth_override.reshape(n_views, 1, hw) if depth_override is not None else \
                noisy.reshape(n_views, 5, hw)[:, 4:5] / scal[0]
            s = mean + scal[1] * depth_eps.reshape(n_views, D, hw)
            z = torch.clip((s + 1.0) / 2.0, 0.0, 1.0) * depth_scale + depth_shift
            zdepth.reshape(-1)[: z.numel()].copy_(z.reshape(-1))
        return self._call(fn)

    def gridattn_tokens(self, feat, zdepth, cams, mask, freqs, ndc_grid, tokens, n_views, S, D, q_first, q_count):
        def fn():
            hw, V = S * S, n_views
            fm = feat.reshape(-1)[: (V + 1) * hw * 256].reshape(V + 1, S, S, 256).float().permute(0, 3, 1, 2)
            z = zdepth.reshape(-1)[: V * D * hw].reshape(V, D, hw)[q_first:q_first + q_count]
            cam = cams.reshape(V + 1, 16)
            R, T, f, pp = cam[:, :9].reshape(-1, 3, 3), cam[:, 9:12], cam[:, 12:14], cam[:, 14:16]
            ctr = -torch.einsum("bj,bij->bi", T, R)
            gx = ndc_grid.reshape(1, S).expand(S, S).reshape(-1)
            gy = ndc_grid.reshape(S, 1).expand(S, S).reshape(-1)
            qi = torch.arange(q_first, q_first + q_count)
            dv = torch.stack([(gx[None] - pp[qi, 0:1]) / f[qi, 0:1], (gy[None] - pp[qi, 1:2]) / f[qi, 1:2],
                              torch.ones(q_count, hw)], -1)
            dirs = dv @ R[qi].transpose(1, 2)  # (q, hw, 3)
            X = ctr[qi][:, None, None, :] + z.permute(0, 2, 1)[..., None] * dirs[:, :, None, :]  # (q, hw, D, 3)
            P = q_count * hw * D
            pts = X.reshape(1, P, 3)

            def sample(fmap, ci):
                v = pts @ R[ci] + T[ci][:, None, :]
                x = f[ci][:, None, 0] * v[..., 0] / v[..., 2] + pp[ci][:, None, 0]
                y = f[ci][:, None, 1] * v[..., 1] / v[..., 2] + pp[ci][:, None, 1]
                g = F.grid_sample(fmap, -torch.stack([x, y], -1).unsqueeze(2), align_corners=True, mode="bilinear",
                                  padding_mode="border")
                return g[..., 0].permute(0, 2, 1)  # (n, P, 256)

            vi = torch.arange(V)
            ref = sample(fm[:V], vi)
            inp = sample(fm[V:], torch.tensor([V])).expand(V, -1, -1)

            def harm(x):
                e = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
                return torch.cat([e.sin(), e.cos(), x], -1)

            rd = pts.expand(V, -1, -1) - ctr[:V, None, :]
            rlen = torch.linalg.norm(rd, dim=-1, keepdim=True)
            rdn = rd / rlen.clamp_min(1e-12)
            ref_pl = harm(torch.cat([rdn, torch.cross(ctr[:V, None, :].expand_as(rdn), rdn, dim=-1)], -1))
            qd = dirs / torch.linalg.norm(dirs, dim=-1, keepdim=True).clamp_min(1e-12)
            qd = qd[:, :, None, :].expand(-1, -1, D, -1).reshape(1, P, 3)
            qo = ctr[qi][:, None, None, :].expand(-1, hw, D, -1).reshape(1, P, 3)
            q_pl = harm(torch.cat([qd, torch.cross(qo, qd, dim=-1)], -1)).expand(V, -1, -1)
            q_dep = harm(z.permute(0, 2, 1).reshape(1, P, 1)).expand(V, -1, -1)
            m = mask.reshape(V, 1, 1).expand(-1, P, -1)
            tok = torch.cat([ref, inp, ref_pl, harm(rlen), q_pl, q_dep, m], -1)  # (V, P, 723)
            o = tokens.reshape(-1)[: P * V * 736].reshape(P, V, 736)
            o.zero_()
            o[:, :, :723] = tok.permute(1, 0, 2).half()
        return self._call(fn)

    def view_attention(self, qkv, out, P, V, heads, hd):
        def fn():
            C = heads * hd
            t = qkv.reshape(-1)[: P * V * 3 * C].reshape(P, V, 3, heads, hd).float().permute(2, 0, 3, 1, 4)
            a = ((t[0] * hd ** -0.5) @ t[1].transpose(-1, -2)).softmax(-1) @ t[2]
            out.reshape(-1)[: P * V * C].copy_(a.transpose(1, 2).reshape(-1).half())
        return self._call(fn)

    def dit_fold_gates(self, jobs):
        def fn():
            for W, gate, bias, Wo, bo in jobs:
                Wo.copy_((W.float() * gate.reshape(-1, 1)).half())
                bo.copy_(gate * bias)
        return self._call(fn)

    def gridattn_dit(self, tokens, token_k, w_pre, b_pre, layers, pool_w, pool_b, pooled, R, V, eps, x_out=None):
        """mvd_gridattn_dit_f16: fp16 operands (tokens, LayerNorm output, q / k / v, attention output, fc1 output), fp32 stream"""
        def fn():
            h = lambda t: t.half().float()
            x = F.gelu(tokens[:R, :token_k].float() @ w_pre[:, :token_k].float().t() + b_pre)
            for lay in layers:
                a = h(F.layer_norm(x, (256,), None, None, eps) * (1 + lay["scale_msa"]) + lay["shift_msa"])
                qkv = h(a @ lay["w_qkv"].float().t() + lay["b_qkv"]).reshape(R // V, V, 8, 3, 32)   # head order: q_h | k_h | v_h
                q, k, v = (qkv[:, :, :, i].permute(0, 2, 1, 3) for i in range(3))
                att = h(((q * 32 ** -0.5) @ k.transpose(-1, -2)).softmax(-1) @ v).permute(0, 2, 1, 3).reshape(R, 256)
                x = x + att @ lay["w_proj"].float().t() + lay["b_proj"]
                a = h(F.layer_norm(x, (256,), None, None, eps) * (1 + lay["scale_mlp"]) + lay["shift_mlp"])
                f = h(F.gelu(a @ lay["w_fc1"].float().t() + lay["b_fc1"]))
                x = x + f @ lay["w_fc2"].float().t() + lay["b_fc2"]
            if x_out is not None:
                x_out.reshape(-1)[: R * 256].copy_(x.reshape(-1))
            v3 = x.reshape(R // V, V, 256)
            wt = (v3 @ pool_w.reshape(256, 1) + pool_b.reshape(-1)[0]).softmax(dim=1)
            pooled.reshape(-1)[: (R // V) * 256].copy_((v3 * wt).sum(1).reshape(-1).half())
        return self._call(fn)

    def view_pool(self, x, w, b, out, P, V, C):
        def fn():
            v = x.reshape(-1)[: P * V * C].reshape(P, V, C)
            wt = (v @ w.reshape(C, 1) + b.reshape(-1)[0]).softmax(dim=1)
            out.reshape(-1)[: P * C].copy_((v * wt).sum(1).reshape(-1).half())
        return self._call(fn)

    def frustum_pool(self, inp, out, n_img, S, D, C, factor):
        def fn():
            v = inp.reshape(-1)[: n_img * S * S * D * C].reshape(n_img, S // factor, factor, S // factor, factor, D, C).float()
            o = v.mean(dim=(2, 4))
            out.reshape(-1)[: o.numel()].copy_(o.reshape(-1).half())
        return self._call(fn)

    def pixel_cross_attn(self, q, kv, out, M, D, heads, dhead):
        def fn():
            C = heads * dhead
            Q = q.reshape(-1)[: M * C].reshape(M, heads, 1, dhead).float()
            KV = kv.reshape(-1)[: M * D * 2 * C].reshape(M, D, 2, heads, dhead).float()
            K, V = KV[:, :, 0].permute(0, 2, 1, 3), KV[:, :, 1].permute(0, 2, 1, 3)
            a = ((Q @ K.transpose(-1, -2)) * dhead ** -0.5).softmax(-1) @ V
            out.reshape(-1)[: M * C].copy_(a.reshape(-1).half())
        return self._call(fn)

    # ------------------------------------------------------------------ training (ABI 15): closed forms; tests/test_training.py pins them to torch.autograd
    def layernorm_fwd(self, x, gamma, beta, y, stats, rows, C, eps):
        def fn():
            v = x.reshape(-1)[: rows * C].reshape(rows, C)
            mean = v.mean(dim=1)
            rstd = (v.var(dim=1, unbiased=False) + eps).rsqrt()
            stats.reshape(-1)[: 2 * rows].reshape(rows, 2).copy_(torch.stack([mean, rstd], dim=1))
            y.reshape(-1)[: rows * C].copy_(F.layer_norm(v, (C,), gamma, beta, eps).reshape(-1))
        return self._call(fn)

    def layernorm_bwd(self, dy, x, gamma, stats, dx, dgamma, dbeta, rows, C):
        def fn():
            v = x.reshape(-1)[: rows * C].reshape(rows, C)
            d = dy.reshape(-1)[: rows * C].reshape(rows, C)
            st = stats.reshape(-1)[: 2 * rows].reshape(rows, 2)
            xh = (v - st[:, :1]) * st[:, 1:]                              # the SAVED statistics, as the kernel uses them
            gx = d if gamma is None else d * gamma.reshape(1, C)
            dxv = st[:, 1:] * (gx - gx.mean(dim=1, keepdim=True) - xh * (gx * xh).mean(dim=1, keepdim=True))
            dx.reshape(-1)[: rows * C].copy_(dxv.reshape(-1))
            if dgamma is not None:
                dgamma.reshape(-1)[:C].copy_((d * xh).sum(0))
                dbeta.reshape(-1)[:C].copy_(d.sum(0))
        return self._call(fn)

    def groupnorm_fwd(self, x, gamma, beta, y, stats, ws, n_img, hw, C, eps, silu):
        def fn():
            v = x.reshape(-1)[: n_img * hw * C].reshape(n_img, hw, 32, C // 32)
            mean = v.mean(dim=(1, 3))
            rstd = (v.var(dim=(1, 3), unbiased=False) + eps).rsqrt()
            stats.reshape(-1)[: n_img * 64].reshape(n_img, 32, 2).copy_(torch.stack([mean, rstd], dim=2))
            o = F.group_norm(v.reshape(n_img, hw, C).permute(0, 2, 1), 32, gamma, beta, eps)
            if silu:
                o = F.silu(o)
            y.reshape(-1)[: n_img * hw * C].copy_(o.permute(0, 2, 1).reshape(-1))
        return self._call(fn)

    def groupnorm_bwd(self, dy, x, gamma, beta, stats, dx, dgamma, dbeta, ws, n_img, hw, C, silu):
        def fn():
            cpg = C // 32
            st = stats.reshape(-1)[: n_img * 64].reshape(n_img, 1, 32, 2)
            v = x.reshape(-1)[: n_img * hw * C].reshape(n_img, hw, 32, cpg)
            d = dy.reshape(-1)[: n_img * hw * C].reshape(n_img, hw, 32, cpg)
            ga, be = gamma.reshape(1, 1, 32, cpg), beta.reshape(1, 1, 32, cpg)
            xh = (v - st[..., :1]) * st[..., 1:]
            dz = d
            if silu:
                z = xh * ga + be
                sg = torch.sigmoid(z)
                dz = d * sg * (1 + z * (1 - sg))
            dgamma.reshape(-1)[:C].copy_((dz * xh).sum(dim=(0, 1)).reshape(-1))
            dbeta.reshape(-1)[:C].copy_(dz.sum(dim=(0, 1)).reshape(-1))
            gz = dz * ga
            s1 = gz.mean(dim=(1, 3), keepdim=True)
            s2 = (gz * xh).mean(dim=(1, 3), keepdim=True)
            dx.reshape(-1)[: n_img * hw * C].copy_((st[..., 1:] * (gz - s1 - xh * s2)).reshape(-1))
        return self._call(fn)

    def act_fwd(self, x, y, rows, cols, mode):
        def fn():
            if mode == ACT_GEGLU:
                a, gate = x.reshape(-1)[: rows * 2 * cols].reshape(rows, 2 * cols).chunk(2, dim=1)
                y.reshape(-1)[: rows * cols].copy_((a * F.gelu(gate)).reshape(-1))
            else:
                v = x.reshape(-1)[: rows * cols]
                y.reshape(-1)[: rows * cols].copy_(F.gelu(v) if mode == ACT_GELU else F.silu(v))
        return self._call(fn)

    def act_bwd(self, dy, x, dx, rows, cols, mode):
        def fn():
            n_in = rows * cols * (2 if mode == ACT_GEGLU else 1)
            with torch.enable_grad():
                v = x.detach().reshape(-1)[:n_in].clone().requires_grad_(True)
                if mode == ACT_GEGLU:
                    a, gate = v.reshape(rows, 2 * cols).chunk(2, dim=1)
                    out = (a * F.gelu(gate)).reshape(-1)
                else:
                    out = F.gelu(v) if mode == ACT_GELU else F.silu(v)
                (g,) = torch.autograd.grad(out, v, dy.detach().reshape(-1)[: rows * cols])
            dx.reshape(-1)[:n_in].copy_(g)
        return self._call(fn)

    def _gather_grid(self, xy, V, P):
        return xy.reshape(-1)[: V * P * 2].reshape(V, P, 1, 2)

    def bilinear_gather_fwd(self, fmap, xy, out, V, H, W, C, P):
        def fn():
            m = fmap.reshape(-1)[: V * H * W * C].reshape(V, H, W, C).permute(0, 3, 1, 2)
            g = F.grid_sample(m, self._gather_grid(xy, V, P), mode="bilinear", padding_mode="border", align_corners=True)   # [V, C, P, 1]
            out.reshape(-1)[: V * P * C].copy_(g[..., 0].permute(0, 2, 1).reshape(-1))
        return self._call(fn)

    def bilinear_gather_bwd(self, dout, xy, dfmap, V, H, W, C, P):
        def fn():
            with torch.enable_grad():
                m = torch.zeros(V, C, H, W, requires_grad=True)
                g = F.grid_sample(m, self._gather_grid(xy, V, P).detach(), mode="bilinear", padding_mode="border", align_corners=True)
                (gm,) = torch.autograd.grad(g, m, dout.detach().reshape(-1)[: V * P * C].reshape(V, P, C).permute(0, 2, 1).unsqueeze(-1))
            dfmap.reshape(-1)[: V * H * W * C].copy_(gm.permute(0, 2, 3, 1).reshape(-1))
        return self._call(fn)
