"""mvdfusion/attention.py of the reference: the view-aligned transformer inserted after every SpatialTransformer of the
middle / output blocks.  Parameter containers (reference names, incl. the `DualAttnetionBlock` spelling) + a native forward."""
import torch
import torch.nn as nn

from .sd_modules import CrossAttention, FeedForward, NativeModule, Normalize, nchw_in, nchw_out, zero_module


class DualAttnetionBlock(nn.Module):
    """mvdfusion/attention.py:16-66 (parameter holder; emitted by engine.Builder.view_aligned_transformer)."""

    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, checkpoint=True,
                 disable_self_attn=False, preserve_unet_dim=False):
        super().__init__()
        assert disable_self_attn is False
        self.disable_self_attn = disable_self_attn
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout, context_dim=None)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout,
                                    context_dim=context_dim if not preserve_unet_dim else None)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.checkpoint = checkpoint


class ViewAlignedFeatureTransformer(NativeModule):
    """mvdfusion/attention.py:72-145"""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None, disable_self_attn=False,
                 use_linear=True, use_checkpoint=True, image_size=None):
        super().__init__()
        if not use_linear or depth != 1:
            raise NotImplementedError("hot path: linear projections, depth 1")
        if context_dim is not None and not isinstance(context_dim, (list, tuple)):
            context_dim = [context_dim]
        self.image_size = image_size
        self.in_channels = in_channels
        self.n_heads = n_heads
        inner_dim = n_heads * d_head
        self.aligned_attn_norm = Normalize(in_channels)
        self.aligned_attn_proj_in = nn.Linear(in_channels, inner_dim)
        self.aligned_attn_transformer_blocks = nn.ModuleList(
            [DualAttnetionBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim[d],
                                disable_self_attn=disable_self_attn, checkpoint=use_checkpoint) for d in range(depth)])
        self.aligned_attn_proj_out = zero_module(nn.Linear(in_channels, inner_dim))
        self.use_linear = use_linear
        self.level_mapper = {self.image_size: 0, self.image_size // 2: 1, self.image_size // 4: 2, self.image_size // 8: 3}

    def forward(self, x, context=None):
        """x (B,C,H,W); context: list of frustum-feature levels (B,h,w,D,768) (mvdfusion/unet.py:198-209)."""
        n, C, H, Wd = x.shape
        ctx = context[self.level_mapper[H]]
        D, Cc = ctx.shape[3], ctx.shape[4]

        def make(plan, b):
            b.heads = self.n_heads
            rows = nchw_in(b, plan, "x", n, C, H * H)
            c32 = b.ops.empty((n * H * H * D, Cc), torch.float32)
            plan.inputs["ctx"] = c32
            c16 = b.cast16(c32, n * H * H * D, Cc)
            y = b.view_aligned_transformer(rows, "", n, H, C, c16, D)
            nchw_out(b, plan, "y", y, n, C, H * H)

        return self._execute(self._plan(("fwd", n, H, D), make), {"x": x, "ctx": ctx}).reshape(n, C, H, Wd)
