"""Training forward + backward of the multi-view denoiser (BASELINE configs[3]) — first slice.

`ViewFusion.forward(batch, trainer_config)` (mvdfusion/viewfusion_zero_depth_rgb.py:362-397 of the reference: prepare_batch ->
shared-t q_sample -> apply_model(cfg_scale = 1.0, is_train = True) -> MSE(noise, prediction)) must return a loss that
`loss.backward()` can differentiate (train.py:90-95).  The inference path is a flat program of fused kernels over recycled fp16
buffers and cannot be differentiated; training therefore runs this module: the same arithmetic, op by op, as a torch autograd graph
whose CONTRACTIONS — every nn.Linear, 1x1 / 3x3 convolution, i.e. ~94 % of the forward and backward FLOPs — are autograd
Functions over the library's tcgen05 GEMM / implicit-GEMM kernel (mvd_gemm_f16):

    forward   y  = x W^T            fp16 operands, fp32 accumulate (the same kernel and packing as inference)
    dgrad     dx = dy W             the same kernel with the roles swapped: A = dy [M, N], "weights" = W^T [K, N]
    wgrad     dW = dy^T x           A = dy^T [N, M], "weights" = x^T [K, M]: contraction over the M rows
    conv3x3   dgrad = implicit-GEMM convolution of dy with the flipped / transposed kernel; wgrad = dy^T im2col(x)

Second slice (ABI 15, csrc/train.cu): the normalisations and activations between the GEMMs — nn.LayerNorm and the adaLN-modulated
LayerNorm of the DiT blocks, GroupNorm32 (+ SiLU), GELU, SiLU, GEGLU — are autograd Functions over the library's fp32 forward /
backward kernels (forward saves per-row / per-group (mean, rstd); backward recomputes the normalised values from them).
MVD_TRAIN_ATEN_POINTWISE=1 puts the ATen ops back for an A/B.

GridAttn's bilinear gather and its scatter-add backward run on mvd_bilinear_gather_{fwd,bwd}_f32 (ABI 16).

What is NOT native yet: the softmax / attention cores (torch SDPA) and the data movement (concat, pad, pooling) run as ATen ops
inside the same autograd graph.  Gradients are fp32; parameters stay the nn.Module's fp32
tensors, so torch optimizers and DistributedDataParallel (train.py:38,93-95) work unchanged.
"""
import math
import os

import torch
import torch.nn.functional as F

from . import runtime
from .ops import GN_WS_BYTES_PER_IMAGE

Z_SCALE = 0.18215


# ------------------------------------------------------------------------------------------------ contraction primitives
def _stream(t):
    return runtime.current_stream(t.device)


def _pad_last(t, mult=8):
    k = t.shape[-1]
    kp = (k + mult - 1) // mult * mult
    return t if kp == k else F.pad(t, (0, kp - k))


def gemm_nt(a16, b16, bias=None):
    """fp32 [M, N] = a16 [M, K] @ b16 [N, K]^T (+ bias) on the tcgen05 GEMM kernel; K is zero-padded to a multiple of 8."""
    a16, b16 = _pad_last(a16.contiguous()), _pad_last(b16.contiguous())
    M, K = a16.shape
    N = b16.shape[0]
    ops = runtime.get_ops(a16.device)
    out = ops.empty((M, N), torch.float32)
    ops.gemm(a16, b16, out, M, N, K, bias=bias)(_stream(a16))
    return out


def _scaled_half(t):
    """fp16 copy of a gradient tensor behind a per-tensor power-of-two scale (exact), chosen on the device so that the largest
    magnitude lands near 2^14: back-propagated gradients are ~1e-7 here and would flush to zero in fp16 (the reference trains in
    fp32; BASELINE configs[3] asks bf16 — the kernel's operand type is fp16, so the range is handled by scaling instead).
    Returns (t * s as fp16, 1 / s as a 0-d fp32 tensor)."""
    amax = t.detach().abs().amax().float().clamp_min(1e-30)
    s = torch.exp2(torch.floor(torch.log2(16384.0 / amax)))
    return (t * s).half(), 1.0 / s


class _LinearFn(torch.autograd.Function):
    """y [M, N] = x [M, K] W[N, K]^T + b"""

    @staticmethod
    def forward(ctx, x, w, b):
        x16, w16 = x.half(), w.half()
        ctx.save_for_backward(x16, w16)
        ctx.has_bias = b is not None
        return gemm_nt(x16, w16, None if b is None else b.float().contiguous())

    @staticmethod
    def backward(ctx, dy):
        x16, w16 = ctx.saved_tensors
        dy16, inv = _scaled_half(dy)
        dx = gemm_nt(dy16, w16.t()) * inv if ctx.needs_input_grad[0] else None     # dy [M, N] x (W^T) [K, N]
        dw = gemm_nt(dy16.t(), x16.t()) * inv if ctx.needs_input_grad[1] else None # dy^T [N, M] x (x^T) [K, M]
        db = dy.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db


def linear(x, w, b=None):
    """nn.Linear / 1x1 convolution on channels-last rows: x [..., K] -> [..., N]"""
    lead = x.shape[:-1]
    y = _LinearFn.apply(x.reshape(-1, x.shape[-1]), w.reshape(w.shape[0], -1), b)
    return y.reshape(*lead, w.shape[0])


def _im2col_t(x16):
    """x16 [n, H, W, C] -> the transposed im2col matrix [9C, n*H*W] (row (ky*3+kx)*C + c), zero padding 1"""
    n, H, W, C = x16.shape
    xp = F.pad(x16, (0, 0, 1, 1, 1, 1))
    taps = torch.stack([xp[:, ky:ky + H, kx:kx + W, :] for ky in range(3) for kx in range(3)])  # [9, n, H, W, C]
    return taps.reshape(9, n * H * W, C).permute(0, 2, 1).reshape(9 * C, n * H * W)


class _Conv3x3Fn(torch.autograd.Function):
    """3x3 convolution, stride 1, zero padding 1, channels-last: x [n, H, W, Cin] (W a power of two), w [Cout, Cin, 3, 3]"""

    @staticmethod
    def forward(ctx, x, w, b):
        n, H, W, Cin = x.shape
        Cout = w.shape[0]
        x16 = _pad_last(x.half())                                                   # channels to a multiple of 8 (the stem has 10)
        Cp = x16.shape[-1]
        w16 = F.pad(w.permute(0, 2, 3, 1), (0, Cp - Cin)).reshape(Cout, 9 * Cp).half().contiguous()  # k = (ky, kx, c)
        ops = runtime.get_ops(x.device)
        out = ops.empty((n * H * W, Cout), torch.float32)
        ops.gemm(x16.contiguous(), w16, out, n * H * W, Cout, 9 * Cp, conv=(n, H, W, Cp), bias=None if b is None else b.float().contiguous())(_stream(x))
        ctx.save_for_backward(x16, w)
        ctx.dims = (n, H, W, Cin, Cout, Cp)
        ctx.has_bias = b is not None
        return out.reshape(n, H, W, Cout)

    @staticmethod
    def backward(ctx, dy):
        x16, w = ctx.saved_tensors
        n, H, W, Cin, Cout, Cp = ctx.dims
        M = n * H * W
        dy16, inv = _scaled_half(dy)
        dy16 = _pad_last(dy16.contiguous())                                          # [n, H, W, Coutp]
        Cop = dy16.shape[-1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dgrad: the same implicit-GEMM convolution on dy with the kernel flipped in space and transposed in the channels
            wf = F.pad(w.flip(2, 3).permute(1, 2, 3, 0), (0, Cop - Cout)).reshape(Cin, 9 * Cop).half().contiguous()
            ops = runtime.get_ops(dy.device)
            out = ops.empty((M, Cin), torch.float32)
            ops.gemm(dy16, wf, out, M, Cin, 9 * Cop, conv=(n, H, W, Cop))(_stream(dy))
            dx = out.reshape(n, H, W, Cin) * inv
        if ctx.needs_input_grad[1]:
            # wgrad: dW[co, (ky, kx, c)] = sum over pixels dy[m, co] * patch[m, (ky, kx, c)] — a GEMM contracting the M pixels
            dwf = gemm_nt(dy16.reshape(M, Cop)[:, :Cout].t(), _im2col_t(x16))         # [Cout, 9 Cp]
            dw = dwf.reshape(Cout, 3, 3, Cp)[..., :Cin].permute(0, 3, 1, 2) * inv
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum((0, 1, 2))
        return dx, dw, db


def conv3x3(x, w, b=None):
    return _Conv3x3Fn.apply(x, w, b)


def conv3x3_stride2(x, w, b=None):
    """Downsample.op (openaimodel.py:151): conv3x3, stride 2, padding 1 = strided patches (torch data movement, differentiable) +
    the GEMM Function"""
    n, H, W, C = x.shape
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    cols = torch.cat([xp[:, ky:ky + H:2, kx:kx + W:2, :] for ky in range(3) for kx in range(3)], dim=-1)  # [n, H/2, W/2, 9C]
    return linear(cols, w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), b)


# ------------------------------------------------------------------------------------------------ normalisation / activation primitives
ATEN_POINTWISE = os.environ.get("MVD_TRAIN_ATEN_POINTWISE") == "1"   # A/B: the first slice's ATen ops instead of csrc/train.cu
_ACT_GELU, _ACT_SILU, _ACT_GEGLU = 1, 2, 3                           # mvd_act_{fwd,bwd}_f32 modes (include/mvd_b200.h)


def _f32c(t):
    return t.detach().float().contiguous()


class _LayerNormFn(torch.autograd.Function):
    """y = LayerNorm(x) * gamma + beta over the last dimension; gamma / beta [C] or both None (mvd_layernorm_{fwd,bwd}_f32)"""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        C = x.shape[-1]
        x2 = _f32c(x).reshape(-1, C)
        rows = x2.shape[0]
        ops = runtime.get_ops(x.device)
        g = None if gamma is None else _f32c(gamma).reshape(C)
        b = None if beta is None else _f32c(beta).reshape(C)
        y, stats = ops.empty((rows, C), torch.float32), ops.empty((rows, 2), torch.float32)
        ops.layernorm_fwd(x2, g, b, y, stats, rows, C, float(eps))(_stream(x))
        ctx.save_for_backward(x2, g, stats)
        ctx.affine = gamma is not None
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, g, stats = ctx.saved_tensors
        rows, C = x2.shape
        ops = runtime.get_ops(dy.device)
        d = _f32c(dy).reshape(rows, C)
        dx = ops.empty((rows, C), torch.float32)
        dg = ops.empty((C,), torch.float32) if ctx.affine else None
        db = ops.empty((C,), torch.float32) if ctx.affine else None
        ops.layernorm_bwd(d, x2, g, stats, dx, dg, db, rows, C)(_stream(dy))
        return dx.reshape(dy.shape), dg, db, None


def layer_norm(x, gamma, beta, eps):
    if ATEN_POINTWISE:
        return F.layer_norm(x, (x.shape[-1],), gamma, beta, eps)
    return _LayerNormFn.apply(x, gamma, beta, eps)


class _GroupNormFn(torch.autograd.Function):
    """GroupNorm32 (+ SiLU) on channels-last x [n, ..., C] (mvd_groupnorm_{fwd,bwd}_f32)"""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, silu):
        n, C = x.shape[0], x.shape[-1]
        xc = _f32c(x)
        hw = xc.numel() // (n * C)
        ops = runtime.get_ops(x.device)
        g, b = _f32c(gamma), _f32c(beta)
        y, stats = ops.empty(tuple(x.shape), torch.float32), ops.empty((n, 32, 2), torch.float32)
        ws = ops.empty((n * GN_WS_BYTES_PER_IMAGE,), torch.uint8)
        ops.groupnorm_fwd(xc, g, b, y, stats, ws, n, hw, C, float(eps), silu)(_stream(x))
        ctx.save_for_backward(xc, g, b, stats)
        ctx.dims = (n, hw, C, bool(silu))
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, g, b, stats = ctx.saved_tensors
        n, hw, C, silu = ctx.dims
        ops = runtime.get_ops(dy.device)
        dx = ops.empty(tuple(xc.shape), torch.float32)
        dgb = ops.empty((2, C), torch.float32)                 # (dgamma | dbeta) as one vector: one memset in the call
        dg, db = dgb[0], dgb[1]
        ws = ops.empty((n * GN_WS_BYTES_PER_IMAGE,), torch.uint8)
        ops.groupnorm_bwd(_f32c(dy), xc, g, b, stats, dx, dg, db, ws, n, hw, C, silu)(_stream(dy))
        return dx, dg, db, None, None


class _ActFn(torch.autograd.Function):
    """GELU (exact erf) / SiLU elementwise, GEGLU over the last dimension ([..., 2 I] -> [..., I]) (mvd_act_{fwd,bwd}_f32)"""

    @staticmethod
    def forward(ctx, x, mode):
        xc = _f32c(x)
        ops = runtime.get_ops(x.device)
        if mode == _ACT_GEGLU:
            cols = x.shape[-1] // 2
            rows = xc.numel() // (2 * cols)
            y = ops.empty(tuple(x.shape[:-1]) + (cols,), torch.float32)
        else:
            rows, cols = 1, xc.numel()
            if cols >= 1 << 31:                          # cols is an int32 in the ABI
                cols = x.shape[-1]
                rows = xc.numel() // cols
            y = ops.empty(tuple(x.shape), torch.float32)
        ops.act_fwd(xc, y, rows, cols, mode)(_stream(x))
        ctx.save_for_backward(xc)
        ctx.dims = (rows, cols, mode)
        return y

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        rows, cols, mode = ctx.dims
        ops = runtime.get_ops(dy.device)
        dx = ops.empty(tuple(xc.shape), torch.float32)
        ops.act_bwd(_f32c(dy), xc, dx, rows, cols, mode)(_stream(dy))
        return dx, None


class _GatherFn(torch.autograd.Function):
    """grid_sample(bilinear, border, align_corners=True) of a channels-last map at fixed coordinates: fmap [V, H, W, C], xy [V, P, 2]
    -> [V, P, C]; the backward scatters dout into the four taps (mvd_bilinear_gather_{fwd,bwd}_f32)"""

    @staticmethod
    def forward(ctx, fmap, xy):
        V, H, W, C = fmap.shape
        P = xy.shape[1]
        ops = runtime.get_ops(fmap.device)
        f, g = _f32c(fmap), _f32c(xy)
        out = ops.empty((V, P, C), torch.float32)
        ops.bilinear_gather_fwd(f, g, out, V, H, W, C, P)(_stream(fmap))
        ctx.save_for_backward(g)
        ctx.dims = (V, H, W, C, P)
        return out

    @staticmethod
    def backward(ctx, dout):
        (g,) = ctx.saved_tensors
        V, H, W, C, P = ctx.dims
        ops = runtime.get_ops(dout.device)
        dfmap = ops.empty((V, H, W, C), torch.float32)
        ops.bilinear_gather_bwd(_f32c(dout), g, dfmap, V, H, W, C, P)(_stream(dout))
        return dfmap, None


def bilinear_gather(fmap, xy):
    assert not xy.requires_grad, "the sampling coordinates carry no gradient on this path"
    return _GatherFn.apply(fmap, xy)


def gelu(x):
    return F.gelu(x) if ATEN_POINTWISE else _ActFn.apply(x, _ACT_GELU)


def silu(x):
    return F.silu(x) if ATEN_POINTWISE else _ActFn.apply(x, _ACT_SILU)


def geglu(h):
    """external/sd1/ldm/modules/attention.py:42-44: h = proj(x) = (a | gate) -> a * gelu(gate)"""
    if ATEN_POINTWISE:
        a, gate = h.chunk(2, dim=-1)
        return a * F.gelu(gate)
    return _ActFn.apply(h, _ACT_GEGLU)


# ------------------------------------------------------------------------------------------------ UNet (channels-last activations)
def _gn(P, p, x, eps, silu=False):
    if ATEN_POINTWISE:
        y = F.group_norm(x.permute(0, 3, 1, 2), 32, P[p + ".weight"], P[p + ".bias"], eps).permute(0, 2, 3, 1)
        return F.silu(y) if silu else y
    return _GroupNormFn.apply(x, P[p + ".weight"], P[p + ".bias"], eps, silu)


def _ln(P, p, x):
    return layer_norm(x, P[p + ".weight"], P[p + ".bias"], 1e-5)


def _lin(P, p, x):
    return linear(x, P[p + ".weight"], P.get(p + ".bias"))


def _attention(q, k, v, heads):
    """softmax(q k^T d^-1/2) v per head (external/sd1/ldm/modules/attention.py:176-190); q [b, n, C], k / v [b, m, C]"""
    b, n, c = q.shape
    d = c // heads
    split = lambda t: t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3)
    o = F.scaled_dot_product_attention(split(q), split(k), split(v))
    return o.permute(0, 2, 1, 3).reshape(b, n, c)


def cross_attention(P, p, x, context, heads):
    ctx = x if context is None else context
    q = linear(x, P[p + ".to_q.weight"])
    k = linear(ctx, P[p + ".to_k.weight"])
    v = linear(ctx, P[p + ".to_v.weight"])
    return _lin(P, p + ".to_out.0", _attention(q, k, v, heads))


def feed_forward(P, p, x):
    return _lin(P, p + ".net.2", geglu(_lin(P, p + ".net.0.proj", x)))


def spatial_transformer(P, p, x, context, heads):
    """external/sd1/ldm/modules/attention.py:268-287 + :219-223"""
    n, H, W, C = x.shape
    h = linear(_gn(P, p + ".norm", x, 1e-6), P[p + ".proj_in.weight"], P[p + ".proj_in.bias"]).reshape(n, H * W, C)
    tb = p + ".transformer_blocks.0"
    h = cross_attention(P, tb + ".attn1", _ln(P, tb + ".norm1", h), None, heads) + h
    h = cross_attention(P, tb + ".attn2", _ln(P, tb + ".norm2", h), context, heads) + h
    h = feed_forward(P, tb + ".ff", _ln(P, tb + ".norm3", h)) + h
    return linear(h.reshape(n, H, W, C), P[p + ".proj_out.weight"], P[p + ".proj_out.bias"]) + x


def view_aligned_transformer(P, p, x, pyramid, heads, image_size):
    """mvdfusion/attention.py:119-145 + :43-66; pyramid[level]: (n, h, w, D, 768)"""
    n, H, W, C = x.shape
    ctx = pyramid[{image_size: 0, image_size // 2: 1, image_size // 4: 2, image_size // 8: 3}[H]]
    h = _lin(P, p + ".aligned_attn_proj_in", _gn(P, p + ".aligned_attn_norm", x, 1e-6).reshape(n, H * W, C))
    tb = p + ".aligned_attn_transformer_blocks.0"
    h = cross_attention(P, tb + ".attn1", _ln(P, tb + ".norm1", h), None, heads) + h
    hp = h.reshape(n * H * W, 1, C)
    cp = ctx.reshape(n * H * W, ctx.shape[3], ctx.shape[4])
    hp = cross_attention(P, tb + ".attn2", _ln(P, tb + ".norm2", hp), cp, heads) + hp
    h = hp.reshape(n, H * W, C)
    h = feed_forward(P, tb + ".ff", _ln(P, tb + ".norm3", h)) + h
    return _lin(P, p + ".aligned_attn_proj_out", h).reshape(n, H, W, C) + x


def resblock(P, p, x, emb):
    """external/sd1/ldm/modules/diffusionmodules/openaimodel.py:255-275"""
    h = conv3x3(_gn(P, p + ".in_layers.0", x, 1e-5, silu=True), P[p + ".in_layers.2.weight"], P[p + ".in_layers.2.bias"])
    h = h + _lin(P, p + ".emb_layers.1", silu(emb))[:, None, None, :]
    h = conv3x3(_gn(P, p + ".out_layers.0", h, 1e-5, silu=True), P[p + ".out_layers.3.weight"], P[p + ".out_layers.3.bias"])
    if p + ".skip_connection.weight" in P:
        x = linear(x, P[p + ".skip_connection.weight"], P[p + ".skip_connection.bias"])
    return x + h


def _run_block(P, p, x, emb, context, pyramid, heads, image_size):
    j = 0
    while True:
        q = f"{p}.{j}"
        if q + ".in_layers.0.weight" in P:
            x = resblock(P, q, x, emb)
        elif q + ".aligned_attn_norm.weight" in P:
            x = view_aligned_transformer(P, q, x, pyramid, heads, image_size)
        elif q + ".norm.weight" in P:
            x = spatial_transformer(P, q, x, context, heads)
        elif q + ".op.weight" in P:
            x = conv3x3_stride2(x, P[q + ".op.weight"], P[q + ".op.bias"])
        elif q + ".conv.weight" in P:
            x = conv3x3(x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2), P[q + ".conv.weight"], P[q + ".conv.bias"])
        elif q + ".weight" in P:
            x = conv3x3(x, P[q + ".weight"], P[q + ".bias"])
        else:
            return x
        j += 1


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def unet_forward(P, p, x, timesteps, context, pyramid, *, model_channels, num_heads, image_size):
    """mvdfusion/unet.py:524-556; x (n, S, S, C_in) channels-last -> (n, S, S, C_out)"""
    emb = _lin(P, p + "time_embed.2", silu(_lin(P, p + "time_embed.0", timestep_embedding(timesteps, model_channels))))

    def count(name):
        pre = f"{p}{name}."
        return 1 + max(int(k[len(pre):].split(".")[0]) for k in P if k.startswith(pre))

    hs, h = [], x
    for i in range(count("input_blocks")):
        h = _run_block(P, f"{p}input_blocks.{i}", h, emb, context, pyramid, num_heads, image_size)
        hs.append(h)
    h = _run_block(P, f"{p}middle_block", h, emb, context, pyramid, num_heads, image_size)
    for i in range(count("output_blocks")):
        h = _run_block(P, f"{p}output_blocks.{i}", torch.cat([h, hs.pop()], dim=-1), emb, context, pyramid, num_heads, image_size)
    return conv3x3(_gn(P, p + "out.0", h, 1e-5, silu=True), P[p + "out.2.weight"], P[p + "out.2.bias"])


def volume_pyramid(vol, num_levels):
    """mvdfusion/unet.py:198-209: area pooling of (n, S, S, D, C) by 1, 2, 4, 8"""
    n, S, _, D, C = vol.shape
    v = vol.permute(0, 3, 4, 1, 2).reshape(n * D, C, S, S)
    out = []
    for i in range(num_levels):
        lv = v if i == 0 else F.avg_pool2d(v, 1 << i)
        hh = lv.shape[-1]
        out.append(lv.reshape(n, D, C, hh, hh).permute(0, 3, 4, 1, 2))
    return out


# ------------------------------------------------------------------------------------------------ GridAttn
def _cam_center(c):
    return -torch.einsum("bj,bij->bi", c["T"], c["R"])


def _project_xy(c, pts):
    v = pts @ c["R"] + c["T"][:, None, :]
    x = c["f"][:, None, 0] * v[..., 0] / v[..., 2] + c["p"][:, None, 0]
    y = c["f"][:, None, 1] * v[..., 1] / v[..., 2] + c["p"][:, None, 1]
    return torch.stack([x, y], dim=-1)


def _unproject(c, xy, depth):
    X = (xy[..., 0] - c["p"][:, None, 0]) * depth / c["f"][:, None, 0]
    Y = (xy[..., 1] - c["p"][:, None, 1]) * depth / c["f"][:, None, 1]
    v = torch.stack([X, Y, torch.full_like(X, depth)], dim=-1)
    return (v - c["T"][:, None, :]) @ c["R"].transpose(1, 2)


def _harmonic(x, n=7, omega0=0.1):
    freqs = (2.0 ** torch.arange(n, dtype=torch.float32, device=x.device)) * omega0
    e = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
    return torch.cat((e.sin(), e.cos(), x), dim=-1)


def _plucker(o, d):
    return _harmonic(torch.cat((d, torch.cross(o.expand_as(d), d, dim=-1)), dim=-1))


def gridattn_forward(P, p, noisy, cams, in_cams, t_embed, sac, somac, depth_eps, input_latents, *, D, num_heads=8, depth_scale=2.0,
                     depth_shift=0.5):
    """GridAttn.forward + aggregate_features (mvdfusion/view_attn_efficient2.py:269-442) -> (N, S, S, D, 768).  sac / somac =
    sqrt(alphas_cumprod)[t], sqrt(1 - alphas_cumprod)[t] (scalars or 0-d tensors); depth_eps (N, D, S, S) stands in for torch.normal's draw."""
    N, _, S, _ = noisy.shape
    dev = noisy.device
    depth = (noisy[:, 4:] / sac).expand(-1, D, -1, -1) + (somac / sac / 10.0) * depth_eps
    zdepth = (torch.clip((depth + 1.0) / 2.0, 0.0, 1.0) * depth_scale + depth_shift).detach()   # torch.normal: no gradient path
    zemb = lambda x: gelu(linear(x.permute(0, 2, 3, 1), P[p + "z_embedder.0.weight"], P[p + "z_embedder.0.bias"]))   # channels-last (V, S, S, 256)
    feat, in_feat = zemb(noisy), zemb(input_latents)
    half = 1.0 / float(S)
    lin_ = torch.linspace(1.0 - half, -1.0 + half, S, dtype=torch.float32, device=dev)
    gy, gx = torch.meshgrid(lin_, lin_, indexing="ij")
    xy = torch.stack([gx, gy], dim=-1).reshape(1, S * S, 2).expand(N, -1, -1)
    p1, p2 = _unproject(cams, xy, 1.0), _unproject(cams, xy, 2.0)
    d = (p2 - p1).reshape(N, S, S, 3)
    o = (p1 - (p2 - p1)).reshape(N, S, S, 3)
    lengths = zdepth.permute(0, 2, 3, 1)                                            # (N, S, S, D)
    xyz = o[..., None, :] + lengths[..., :, None] * d[..., None, :]
    pts = xyz.reshape(1, N * S * S * D, 3)
    V, HWD = N, S * S * D

    def sample(fmap, c):
        xy = -_project_xy(c, pts)                                                                  # (V', P, 2), no gradient path (zdepth is detached)
        if ATEN_POINTWISE:
            g = F.grid_sample(fmap.permute(0, 3, 1, 2), xy.unsqueeze(2), align_corners=True, mode="bilinear", padding_mode="border")
            return g[..., 0].reshape(g.shape[0], g.shape[1], N, HWD).permute(0, 2, 3, 1)
        return bilinear_gather(fmap, xy).reshape(fmap.shape[0], N, HWD, fmap.shape[-1])

    ref_feat = sample(feat, cams)
    inp_feat = sample(in_feat, in_cams).expand(V, -1, -1, -1)
    centers = _cam_center(cams)
    ref_dir = (pts.expand(V, -1, -1) - centers[:, None, :]).reshape(V, N, HWD, 3)
    ref_depth = _harmonic(torch.linalg.norm(ref_dir, dim=-1, keepdim=True))
    ref_pl = _plucker(centers[:, None, None, :], F.normalize(ref_dir, dim=-1))
    q_dir = F.normalize(d, dim=-1).reshape(1, N, S * S, 1, 3).expand(1, N, S * S, D, 3).reshape(1, N, HWD, 3)
    q_pl = _plucker(centers[None, :, None, :], q_dir).expand(V, -1, -1, -1)
    q_depth = _harmonic(lengths.reshape(1, N, HWD, 1)).expand(V, -1, -1, -1)
    mask = torch.ones(V, N, HWD, 1, device=dev)
    z = torch.cat((ref_feat, inp_feat, ref_pl, ref_depth, q_pl, q_depth, mask), dim=-1)          # (V, N, HWD, 723)
    x = z.reshape(V, N * HWD, z.shape[-1]).permute(1, 0, 2)                                        # (P, V, 723)
    x = gelu(_lin(P, p + "pre_layer_b.0", x))
    c = t_embed[:1]
    C = x.shape[-1]
    hd = C // num_heads
    i = 0
    while f"{p}aggregation_transformer.layer_list.{i}.attn.qkv.weight" in P:
        q = f"{p}aggregation_transformer.layer_list.{i}"
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = _lin(P, q + ".adaLN_modulation.1", silu(c)).chunk(6, dim=1)
        # modulate(norm(x), shift, scale) with one conditioning row (c = t_embed[:1]) is a LayerNorm with gamma = 1 + scale, beta = shift
        modln = lambda v, sh, sc: layer_norm(v, (1 + sc).reshape(C), sh.reshape(C), 1e-6)
        qkv = _lin(P, q + ".attn.qkv", modln(x, sh_a, sc_a)).reshape(x.shape[0], V, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(x.shape[0], V, C)
        x = x + g_a.unsqueeze(1) * _lin(P, q + ".attn.proj", a)
        x = x + g_m.unsqueeze(1) * _lin(P, q + ".mlp.fc2", gelu(_lin(P, q + ".mlp.fc1", modln(x, sh_m, sc_m))))
        i += 1
    w = _lin(P, p + "aggregation_transformer.weight_layer", x).softmax(dim=-2)
    return _lin(P, p + "final_layer_b", (x * w).sum(dim=-2)).reshape(N, S, S, D, -1)


# ------------------------------------------------------------------------------------------------ ViewFusion.apply_model (training)
def named_tensors(model):
    """name -> Parameter / buffer of the facade, the reference's state-dict names (the autograd leaves ARE the module's parameters)"""
    P = dict(model.named_parameters())
    P.update({k: v for k, v in model.named_buffers() if k not in P})
    return P


def apply_model_train(model, noisy, cams, input_latents, in_cams, clip_v_embed, t, *, depth_eps=None, drop_random=None):
    """ViewFusion.apply_model on the cfg_scale == 1.0 branch with is_train=True (viewfusion_zero_depth_rgb.py:282-331; unet.py:129-164),
    differentiable.  cams / in_cams: objects with .R .T .focal_length .principal_point.  Returns the predicted noise (N, 5, S, S)."""
    P = named_tensors(model)
    N, _, S, _ = noisy.shape
    D = model.view_attn.n_pts_per_ray
    um = model.unet_model.unet_model
    cd = lambda c: {"R": c.R.float(), "T": c.T.float(), "f": c.focal_length.float(), "p": c.principal_point.float()}
    sch = model.scheduler
    # the shared timestep stays on the device (no host round trip, graph-capturable).  Indexed with a 1-element vector: a 0-d index
    # tensor would be read back as a Python integer (an implicit .item()), which is what stopped the first capture attempt
    t0 = t.reshape(-1)[:1]
    sac, somac = sch.sqrt_alphas_cumprod[t0].reshape(()), sch.sqrt_one_minus_alphas_cumprod[t0].reshape(())   # 0-d tensors
    if depth_eps is None:
        depth_eps = torch.randn(N, D, S, S, device=noisy.device)
    te = timestep_embedding(t.float(), 256)
    t_embed = _lin(P, "time_embed.2", silu(_lin(P, "time_embed.0", te)))
    vol = gridattn_forward(P, "view_attn.", noisy, cd(cams), cd(in_cams), t_embed, sac, somac, depth_eps, input_latents, D=D,
                           num_heads=model.view_attn.num_heads, depth_scale=model.view_attn.depth_scale, depth_shift=model.view_attn.depth_shift)
    clip_embed = clip_v_embed
    for n_, i in enumerate((0, 2, 4)):
        clip_embed = _lin(P, f"cc_projection.{i}", clip_embed)
        if n_ < 2:
            clip_embed = silu(clip_embed)
    x_concat = input_latents.expand(N, -1, -1, -1)
    if model.drop_conditions:  # unet.py:118-127,140-151
        r = drop_random if drop_random is not None else torch.rand(N, device=noisy.device)
        r = r.to(noisy.device)
        keep = lambda m: 1.0 - m.float()
        drop_all = r <= 0.05
        clip_embed = keep(((r > 0.15) & (r <= 0.2)) | drop_all).view(N, 1, 1) * clip_embed
        vol = keep(((r > 0.1) & (r <= 0.15)) | drop_all).view(N, 1, 1, 1, 1) * vol
        x_concat = keep(((r > 0.05) & (r <= 0.1)) | drop_all).view(N, 1, 1, 1) * x_concat
    xc = torch.cat([x_concat[:, :4] / Z_SCALE, x_concat[:, 4:]], dim=1)
    xin = torch.cat([noisy, xc], dim=1).permute(0, 2, 3, 1)                                        # channels-last
    pyr = volume_pyramid(vol, len(um.channel_mult))
    eps = unet_forward(P, "unet_model.unet_model.", xin, t[:1], clip_embed, pyr, model_channels=um.model_channels, num_heads=um.num_heads,
                       image_size=um.image_size)
    return eps.permute(0, 3, 1, 2)
