"""Parameter containers + native forwards for the Stable-Diffusion-v1 building blocks the reference UNet uses.

Class names, constructor arguments, sub-module / parameter names and initialisation follow
external/sd1/ldm/modules/{attention.py, diffusionmodules/openaimodel.py, diffusionmodules/util.py} of the reference so
that its checkpoints load unchanged (SURVEY.md §8b state-dict contract).  The arithmetic does NOT live here: every
`forward` compiles (once per input shape) a program of sm_100a kernel calls through mvdfusion_b200.engine and replays it.
"""
import torch
import torch.nn as nn

from .. import engine as E
from .. import runtime
from ..runtime import WeightCache, current_stream


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class GroupNorm32(nn.GroupNorm):
    """util.py:215-217 (parameter holder; the normalisation itself is mvd_groupnorm_f32_f16)."""


def normalization(channels):
    return GroupNorm32(32, channels)


def Normalize(in_channels):
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class ModulePlan:
    def __init__(self):
        self.prog = E.Program()
        self.inputs = {}
        self.outputs = {}


class NativeModule(nn.Module):
    """Base of every module whose forward is a compiled kernel program."""

    def _device(self):
        return next(self.parameters()).device

    def _plan(self, key, make):
        ops = runtime.get_ops(self._device())
        cache = self.__dict__.setdefault("_mvd_cache", WeightCache())
        cache.get(self, ops)
        if key not in cache.plans:
            plan = ModulePlan()
            make(plan, E.Builder(ops, cache.pack(self, ops), program=plan.prog))
            cache.plans[key] = plan
        return cache.plans[key]

    def _execute(self, plan, feeds, out="y"):
        for k, v in feeds.items():
            plan.inputs[k].copy_(v.reshape(plan.inputs[k].shape))
        plan.prog.run(current_stream(self._device()))
        return plan.outputs[out].clone()


def nchw_in(b, plan, name, n, C, hw):
    """register an NCHW fp32 input buffer and emit its conversion to rows x channels"""
    src = b.ops.empty((n, C, hw), torch.float32)
    plan.inputs[name] = src
    rows = b.t32(n * hw, C)
    b.prog.append(b.ops.nchw_to_rows(src, rows, n, C, hw))
    return rows


def nchw_out(b, plan, name, rows, n, C, hw, ld=None):
    dst = b.ops.empty((n, C, hw), torch.float32)
    b.prog.append(b.ops.rows_to_nchw(rows, dst, n, C, ld if ld is not None else C, hw))
    plan.outputs[name] = dst
    return dst


# ------------------------------------------------------------------------------------------------ attention.py
class CrossAttention(NativeModule):
    """external/sd1/ldm/modules/attention.py:152-193"""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner_dim = dim_head * heads
        context_dim = context_dim if context_dim is not None else query_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))

    def forward(self, x, context=None, mask=None):
        """Self-attention (context=None) over x (B, seq, C) — the shape class on the hot path.  The one-token and
        per-pixel D-key cross-attentions are fused into their parent blocks (engine.Builder.spatial_transformer /
        view_cross_attention) and are not exposed as a standalone call."""
        if context is not None or mask is not None:
            raise NotImplementedError("standalone CrossAttention.forward supports context=None (self-attention) only")
        n, seq, C = x.shape

        def make(plan, b):
            b.heads = self.heads
            xin = b.ops.empty((n * seq, C), torch.float32)
            plan.inputs["x"] = xin
            a = b.cast16(xin, n * seq, C)
            d = C // self.heads
            dpad = E._round_up(d, 64)
            q, k, vt = b.qkv_buffers(n, seq, dpad)
            b.gemm(a, b.W.qkv(""), q, n * seq, 3 * C, C, qkv=dict(out_k=k, out_vt=vt, heads=self.heads, dhead=d, dpad=dpad, seq=seq))
            b.prog.append(b.ops.attn_self(q, k, vt, a, n, self.heads, seq, d, dpad, C))
            y = b.ops.empty((n * seq, C), torch.float32)
            b.gemm(a, b.W.lin("to_out.0.weight"), y, n * seq, C, C, bias=b.W.f32("to_out.0.bias"))
            plan.outputs["y"] = y

        return self._execute(self._plan(("self", n, seq), make), {"x": x}).reshape(n, seq, C)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """external/sd1/ldm/modules/attention.py:47-64 (parameter holder; fused into Builder.feed_forward)."""

    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.0):
        super().__init__()
        if not glu:
            raise NotImplementedError("the hot path uses gated (GEGLU) feed-forwards only")
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        self.net = nn.Sequential(GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))


class BasicTransformerBlock(nn.Module):
    """external/sd1/ldm/modules/attention.py:195-223 (parameter holder)."""

    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, checkpoint=True,
                 disable_self_attn=False):
        super().__init__()
        if disable_self_attn:
            raise NotImplementedError
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.checkpoint = checkpoint


class SpatialTransformer(NativeModule):
    """external/sd1/ldm/modules/attention.py:225-287"""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None, disable_self_attn=False,
                 use_linear=False, use_checkpoint=True):
        super().__init__()
        if use_linear or depth != 1:
            raise NotImplementedError("hot path: 1x1-conv projections, depth 1")
        if context_dim is not None and not isinstance(context_dim, (list, tuple)):
            context_dim = [context_dim]
        self.in_channels = in_channels
        self.n_heads = n_heads
        inner_dim = n_heads * d_head
        self.norm = Normalize(in_channels)
        self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim[d],
                                   disable_self_attn=disable_self_attn, checkpoint=use_checkpoint) for d in range(depth)])
        self.proj_out = zero_module(nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0))
        self.use_linear = use_linear

    def forward(self, x, context=None):
        """x (B,C,H,W); context (B,1,768): one CLIP token per view, as on the hot path."""
        if isinstance(context, (list, tuple)):
            context = context[0]
        n, C, H, Wd = x.shape
        if context is None or context.shape[1] != 1 or H != Wd:
            raise NotImplementedError("SpatialTransformer.forward: square maps and a single context token per image")

        def make(plan, b):
            b.heads = self.n_heads
            rows = nchw_in(b, plan, "x", n, C, H * H)
            ctx = b.ops.empty((n, context.shape[-1]), torch.float32)
            plan.inputs["ctx"] = ctx
            vec = b.clip_vector(ctx, "", n, C)
            y = b.spatial_transformer(rows, "", n, H, C, vec)
            nchw_out(b, plan, "y", y, n, C, H * H)

        return self._execute(self._plan(("fwd", n, H), make), {"x": x, "ctx": context}).reshape(n, C, H, Wd)


# ------------------------------------------------------------------------------------------------ openaimodel.py
class TimestepBlock(nn.Module):
    pass


class Upsample(nn.Module):
    """openaimodel.py:91-119 (parameter holder; nearest x2 + conv3x3 = Builder.upsample)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if not use_conv or dims != 2:
            raise NotImplementedError
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.conv = nn.Conv2d(self.channels, self.out_channels, 3, padding=padding)


class Downsample(nn.Module):
    """openaimodel.py:134-160 (parameter holder; stride-2 conv3x3 = Builder.downsample)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if not use_conv or dims != 2:
            raise NotImplementedError
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.op = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)


class ResBlock(NativeModule, TimestepBlock):
    """openaimodel.py:163-275"""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if use_conv or use_scale_shift_norm or up or down or dims != 2:
            raise NotImplementedError("hot path: plain ResBlock (1x1 skip, additive timestep embedding)")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(), nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)

    def forward(self, x, emb):
        n, C, H, Wd = x.shape
        ne = emb.shape[0]
        if H != Wd or ne not in (1, n):
            raise NotImplementedError("ResBlock.forward: square maps; emb rows 1 or batch")

        def make(plan, b):
            rows = nchw_in(b, plan, "x", n, C, H * H)
            e = b.ops.empty((ne, self.emb_channels), torch.float32)
            plan.inputs["emb"] = e
            y = b.resblock(rows, "", n, H, C, self.out_channels, e, self.emb_channels)
            nchw_out(b, plan, "y", y, n, self.out_channels, H * H)

        return self._execute(self._plan(("fwd", n, H, ne), make), {"x": x, "emb": emb}).reshape(n, self.out_channels, H, Wd)
