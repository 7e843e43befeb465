"""mvdfusion/unet.py of the reference: the view-conditioned SD-1.x UNet, its CFG / condition-drop wrapper.

Parameter containers with the reference's names (`input_blocks.i.j…`, `middle_block.j…`, `output_blocks.i.j…`, `out.*`,
`time_embed.*`) and native forwards: the whole UNet pass is one compiled program of sm_100a kernel calls
(engine.emit_unet), never a layer-by-layer PyTorch walk.
"""
import os

import torch
import torch.nn as nn

from .. import engine as E
from ..config import instantiate_from_config, load_model_from_config
from ..denoise import UNetStagePlan
from ..runtime import current_stream
from .attention import ViewAlignedFeatureTransformer
from .sd_modules import (Downsample, NativeModule, ResBlock, SpatialTransformer, TimestepBlock, Upsample, normalization,
                         zero_module)


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """mvdfusion/unet.py:36-52 (container; children are dispatched by engine.emit_unet)."""


class UNetModel(NativeModule):
    """mvdfusion/unet.py:215-576"""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False, use_spatial_transformer=True,
                 use_view_aligned_transformer=True, transformer_depth=1, context_dim=None, n_embed=None, legacy=True):
        super().__init__()
        assert use_view_aligned_transformer, "This module only supports spatial view aligned transformer!"
        assert use_spatial_transformer, "This module only supports spatial view aligned transformer"
        assert context_dim is not None, "context_dim (cross-attention conditioning width) is required"
        if isinstance(context_dim, (list, tuple)):
            context_dim = list(context_dim)
        if num_heads == -1 or num_head_channels != -1:
            raise NotImplementedError("hot path: fixed num_heads (num_head_channels == -1)")
        if num_classes is not None or n_embed is not None or resblock_updown or use_scale_shift_norm or dims != 2 or not conv_resample:
            raise NotImplementedError("hot path configuration only (configs/*.yaml of the reference)")
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = tuple(channel_mult)
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.predict_codebook_ids = False
        self.spec = E.UNetSpec(model_channels, channel_mult, num_res_blocks, attention_resolutions, num_heads, image_size,
                               in_channels, out_channels)

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, time_embed_dim), nn.SiLU(),
                                        nn.Linear(time_embed_dim, time_embed_dim))

        def make(layer):
            kind = layer[0]
            if kind == "stem":
                return nn.Conv2d(layer[1], layer[2], 3, padding=1)
            if kind == "res":
                return ResBlock(layer[1], time_embed_dim, dropout, out_channels=layer[2], dims=dims, use_checkpoint=use_checkpoint)
            if kind == "st":
                return SpatialTransformer(layer[1], num_heads, layer[1] // num_heads, depth=transformer_depth, context_dim=context_dim)
            if kind == "vaft":
                return ViewAlignedFeatureTransformer(layer[1], num_heads, layer[1] // num_heads, depth=transformer_depth,
                                                     context_dim=context_dim, image_size=image_size)
            if kind == "down":
                return Downsample(layer[1], conv_resample, dims=dims, out_channels=layer[1])
            if kind == "up":
                return Upsample(layer[1], conv_resample, dims=dims, out_channels=layer[1])
            raise ValueError(kind)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(*[make(l) for l in blk]) for blk in self.spec.input_blocks])
        self.middle_block = TimestepEmbedSequential(*[make(l) for l in self.spec.middle])
        self.output_blocks = nn.ModuleList([TimestepEmbedSequential(*[make(l) for l in blk]) for blk in self.spec.output_blocks])
        self.out = nn.Sequential(normalization(self.spec.final_ch), nn.SiLU(),
                                 zero_module(nn.Conv2d(model_channels, out_channels, 3, padding=1)))

    def forward(self, x, timesteps=None, context=None, y=None, volume_feats=None, **kwargs):
        """x (n, in_channels, S, S); timesteps (1,) or (n,) holding ONE shared value (the path passes t[:1]);
        context (n, 1, 768); volume_feats: list of (n, h, w, D, 768) levels.  Returns (n, out_channels, S, S)."""
        assert y is None, "must specify y if and only if the model is class-conditional"
        n, C, S, S2 = x.shape
        t0 = float(timesteps.reshape(-1)[0])
        if S != S2 or context.shape[1] != 1 or not bool((timesteps == timesteps.reshape(-1)[0]).all()):
            raise NotImplementedError("UNetModel.forward: square latents, one context token, one shared timestep")
        D = volume_feats[0].shape[3]
        spec = self.spec
        # split-precision stem (engine.PackedWeights.conv3_stem_hilo): the input goes in as [x_hi | x_lo | x_hi] channels, each exactly
        # representable in fp16 — boundary glue in torch, the denoising loop does the same inside mvd_unet_input_f16
        stem_hilo = int(os.environ.get("MVD_HILO", "2")) >= 1 and 3 * C <= 32
        if stem_hilo:
            hi = x.float().half().float()
            x = torch.cat([hi, (x.float() - hi).half().float(), hi], dim=1)
            C = 3 * C
        cpad = E._round_up(C, 16)

        def make(plan, b):
            xin = b.ops.empty((n, C, S * S), torch.float32)
            plan.inputs["x"] = xin
            x16 = b.ops.zeros((n * S * S, cpad), torch.float16)
            b.prog.append(b.ops.nchw_to_nhwc16(xin, x16, n, C, S * S, cpad))
            ctx = b.ops.empty((n, E.CTX_DIM), torch.float32)
            plan.inputs["ctx"] = ctx
            t_dev = b.ops.empty((1,), torch.float32)
            plan.inputs["t"] = t_dev
            clipvecs = {p: b.clip_vector(ctx, p, n, Cl) for p, Cl in spec.st_layers()}
            pyr = []
            for l in range(len(spec.mult)):
                h = S >> l
                v32 = b.ops.empty((n * h * h * D, E.CTX_DIM), torch.float32)
                plan.inputs[f"vol{l}"] = v32
                v16 = b.ops.empty((n * h * h * D, E.CTX_DIM), torch.float16)
                b.prog.append(b.ops.cast(v32, v16, v32.numel()))
                pyr.append(v16)
            head = E.emit_unet(b, spec, x16, n, S, D, t_dev, E.timestep_freqs(spec.mc, b.ops.device), clipvecs, pyr, c_in_pad=cpad, stem_hilo=stem_hilo)
            out = b.ops.empty((n, spec.out_channels, S * S), torch.float32)
            b.prog.append(b.ops.rows_to_nchw(head, out, n, spec.out_channels, 8, S * S))
            plan.outputs["y"] = out

        feeds = {"x": x, "ctx": context, "t": torch.tensor([t0])}
        for l, v in enumerate(volume_feats):
            feeds[f"vol{l}"] = v
        return self._execute(self._plan(("fwd", n, S, D, stem_hilo), make), feeds).reshape(n, spec.out_channels, S, S)

    def get_cross_attn_parameters(self, finetune_cross_attn, finetune_view_attn):
        """mvdfusion/unet.py:558-571"""
        out = []
        for name, p in self.named_parameters():
            if finetune_cross_attn and any(s in name for s in (".norm.", ".proj_in.", ".transformer_blocks.", ".proj_out.")):
                out.append(p)
            if finetune_view_attn and ".aligned_attn_" in name:
                out.append(p)
        return out

    def disable_unet_grad(self):
        """mvdfusion/unet.py:573-576"""
        for name, p in self.named_parameters():
            if ".aligned_attn_" not in name:
                p.requires_grad_(False)


class UNetWrapper(nn.Module):
    """mvdfusion/unet.py:56-209: CFG pair, condition drop, zero123 concat scaling, frustum-feature pyramid."""

    PARAM_MAPPER = {  # zero123 layer positions -> positions after the inserted view-aligned transformers (unet.py:70-86)
        "output_blocks.5.2.conv.weight": "output_blocks.5.3.conv.weight",
        "output_blocks.5.2.conv.bias": "output_blocks.5.3.conv.bias",
        "output_blocks.8.2.conv.weight": "output_blocks.8.3.conv.weight",
        "output_blocks.8.2.conv.bias": "output_blocks.8.3.conv.bias",
        **{f"middle_block.2.{k}": f"middle_block.3.{k}" for k in (
            "in_layers.0.weight", "in_layers.0.bias", "in_layers.2.weight", "in_layers.2.bias", "emb_layers.1.weight",
            "emb_layers.1.bias", "out_layers.0.weight", "out_layers.0.bias", "out_layers.3.weight", "out_layers.3.bias")},
    }

    def __init__(self, unet_config, unet_path=None, drop_conditions=False, drop_scheme="default", use_zero_123=False,
                 finetune_unet=False, finetune_cross_attn=False, finetune_view_attn=True, remove_keys=[]):
        super().__init__()
        self.unet_model = load_model_from_config(unet_config, unet_path, verbose=False,
                                                 replace_key=["model.diffusion_model.", ""], ignore_keys=["aligned_attn_"],
                                                 param_mapper=self.PARAM_MAPPER, remove_keys=remove_keys)
        if not finetune_unet:
            self.unet_model.disable_unet_grad()
        self.drop_conditions = drop_conditions
        self.drop_scheme = drop_scheme
        self.use_zero_123 = use_zero_123
        self.finetune_unet = finetune_unet
        self.finetune_cross_attn = finetune_cross_attn
        self.finetune_view_attn = finetune_view_attn

    def get_trainable_parameters(self):
        if self.finetune_unet:
            return self.unet_model.parameters()
        return self.unet_model.get_cross_attn_parameters(finetune_cross_attn=self.finetune_cross_attn,
                                                         finetune_view_attn=self.finetune_view_attn)

    def get_drop_scheme(self, B, device):
        """mvdfusion/unet.py:118-127"""
        if self.drop_scheme != "default":
            raise NotImplementedError
        r = torch.rand(B, dtype=torch.float32, device=device)
        return (r > 0.15) & (r <= 0.2), (r > 0.1) & (r <= 0.15), (r > 0.05) & (r <= 0.1), r <= 0.05

    def _stage(self, n, S, D, use_cfg):
        um = self.unet_model
        from .. import runtime
        from ..runtime import WeightCache
        ops = runtime.get_ops(um._device())
        cache = um.__dict__.setdefault("_mvd_cache", WeightCache())
        cache.get(um, ops)
        key = ("stage", n, S, D, use_cfg)
        if key not in cache.plans:
            cache.plans[key] = UNetStagePlan(ops, cache.pack(um, ops), um.spec, n_views=n, S=S, D=D, use_cfg=use_cfg)
        return cache.plans[key]

    def forward(self, x, t, clip_embed, volume_feats, x_concat=None, is_train=False):
        """x (B,5,S,S); t (1,); clip_embed (B,1,768); volume_feats (B,S,S,D,768); x_concat (B,5,S,S)."""
        if not self.use_zero_123 or x_concat is None:
            raise NotImplementedError("hot path: zero123-style concat conditioning")
        B, _, S, _ = x.shape
        cond_scale = None
        if self.drop_conditions and is_train:
            drop_clip, drop_volume, drop_concat, drop_all = self.get_drop_scheme(B, x.device)
            clip_embed = (1.0 - (drop_clip | drop_all).float()).view(B, 1, 1) * clip_embed
            volume_feats = (1.0 - (drop_volume | drop_all).float()).view(B, 1, 1, 1, 1) * volume_feats
            cond_scale = 1.0 - (drop_concat | drop_all).float()
        plan = self._stage(B, S, volume_feats.shape[3], False)
        return plan.run(x, t.reshape(-1)[0], clip_embed, volume_feats, x_concat, 1.0, current_stream(x.device), cond_scale)

    @torch.no_grad()
    def predict_with_unconditional_scale(self, x, t, clip_embed, volume_feats, x_concat, unconditional_scale):
        B, _, S, _ = x.shape
        plan = self._stage(B, S, volume_feats.shape[3], True)
        return plan.run(x, t.reshape(-1)[0], clip_embed, volume_feats, x_concat, unconditional_scale, current_stream(x.device))

    def get_volume_feats_pyramid(self, volume_feats):
        """mvdfusion/unet.py:198-209 as a standalone call: (b,h,w,d,c) -> list of 'area'-pooled levels."""
        from .. import runtime
        ops = runtime.get_ops(volume_feats.device)
        b, h, w, d, c = volume_feats.shape
        v16 = ops.empty((b * h * w * d, c), torch.float16)
        stream = current_stream(volume_feats.device)
        src = volume_feats.float().contiguous()
        ops.cast(src, v16, src.numel())(stream)
        out = [volume_feats]
        for l in range(1, len(self.unet_model.channel_mult)):
            hl = h >> l
            o16 = ops.empty((b * hl * hl * d, c), torch.float16)
            ops.frustum_pool(v16, o16, b, h, d, c, 1 << l)(stream)
            out.append(o16.float().reshape(b, hl, hl, d, c))
        return out
