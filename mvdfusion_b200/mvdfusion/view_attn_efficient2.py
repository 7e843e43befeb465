"""mvdfusion/view_attn_efficient2.py of the reference: GridAttn, the depth-guided cross-view attention.

Parameter containers carry the reference's names (incl. timm's `attn.qkv / attn.proj / mlp.fc1 / mlp.fc2` inside each DiT
block and the never-called `t_embedder`); the forward is one compiled program (engine.emit_gridattn): depth sampling +
z-embedding, unproject -> reproject -> bilinear gather -> Plücker / depth harmonics, the transformer over the view axis
on tcgen05 GEMMs, view-softmax pooling and the 256 -> 768 projection.
"""
import torch
import torch.nn as nn

from .. import engine as E
from ..denoise import pack_cameras
from .embedder import RayEmbedder, TimestepEmbedder
from .sd_modules import NativeModule


class Attention(nn.Module):
    """timm.models.vision_transformer.Attention parameter layout (qkv with bias, proj)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, **kwargs):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class Mlp(nn.Module):
    """timm.models.vision_transformer.Mlp parameter layout (fc1, exact GELU, fc2)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, **kwargs):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)


class DiTBlock(nn.Module):
    """mvdfusion/view_attn_efficient2.py:42-67 (parameter holder)."""

    def __init__(self, hidden_size, num_heads, cond_dim=None, mlp_ratio=4.0, **block_kwargs):
        super().__init__()
        cond_dim = hidden_size if cond_dim is None else cond_dim
        self.norm1 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.attn = Attention(hidden_size, num_heads=num_heads, qkv_bias=True, **block_kwargs)
        self.norm2 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.mlp = Mlp(in_features=hidden_size, hidden_features=int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(cond_dim, 6 * hidden_size, bias=True))


class AggregationTransformer(nn.Module):
    """mvdfusion/view_attn_efficient2.py:70-93 (parameter holder)."""

    def __init__(self, hidden_size, num_layers=3, num_heads=8, mlp_ratio=2.0, use_t=False):
        super().__init__()
        if not use_t:
            raise NotImplementedError
        self.use_t = use_t
        self.layer_list = nn.ModuleList([DiTBlock(hidden_size, num_heads=num_heads, mlp_ratio=mlp_ratio) for _ in range(num_layers)])
        self.weight_layer = nn.Linear(hidden_size, 1)


class GridAttn(NativeModule):
    """mvdfusion/view_attn_efficient2.py:96-442"""

    def __init__(self, input_size=32, in_channels=4, hidden_size=256, output_dim=768, num_heads=8, mlp_ratio=2.0,
                 num_layers=3, side_length=32, world_scale=0.6, z_near_far_scale=0.8, depth_scale=2.0, depth_shift=0.5,
                 n_pts_per_ray=3, use_t=True, keep_top_k_views=False, top_k=4, device="cpu"):
        super().__init__()
        if keep_top_k_views:
            raise NotImplementedError("keep_top_k_views is never enabled by the reference configs")
        if hidden_size != E.Z_CH or in_channels != 5:
            raise NotImplementedError("hot path: 5 latent channels (4 + depth), hidden size 256")
        self.input_size = input_size
        self.world_scale = world_scale
        self.side_length = side_length
        self.z_near_far_scale = z_near_far_scale
        self.depth_scale = depth_scale
        self.depth_shift = depth_shift
        self.n_pts_per_ray = n_pts_per_ray
        self.keep_top_k_views = keep_top_k_views
        self.top_k = top_k
        self.num_heads = num_heads
        self.num_layers = num_layers
        self.output_dim = output_dim
        n_harmonic = 7
        depth_dim = 1 * (2 * n_harmonic + 1)
        plucker_dim = 6 * (2 * n_harmonic + 1)
        z_output_dim = 256
        self.z_embedder = nn.Sequential(nn.Linear(in_channels, z_output_dim), nn.GELU())
        self.t_embedder = TimestepEmbedder(hidden_size)
        self.ray_embedder = RayEmbedder(input_size)
        self.use_t = use_t
        self.pre_layer_b = nn.Sequential(nn.Linear(z_output_dim * 2 + plucker_dim * 2 + depth_dim * 2 + 1, hidden_size), nn.GELU())
        self.aggregation_transformer = AggregationTransformer(hidden_size=hidden_size, num_layers=num_layers,
                                                              num_heads=num_heads, mlp_ratio=mlp_ratio, use_t=use_t)
        self.final_layer_b = nn.Linear(hidden_size, output_dim)
        for block in self.aggregation_transformer.layer_list:  # adaLN-Zero (:172-176)
            nn.init.constant_(block.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(block.adaLN_modulation[-1].bias, 0)

    def forward(self, noisy_latents, batch_cameras, predict_mask, t_embed, t, scheduler, overwrite_attn_depth=None,
                input_latents=None, input_cameras=None, depth_eps=None):
        """-> (B, S, S, D, output_dim).  `depth_eps` (B, D, S, S): optional standard-normal draws used in place of the
        module's own torch.randn for the depth samples (:431: torch.normal(mean, std) == mean + std * eps)."""
        assert noisy_latents.shape[1] == 5, "depth wise efficient attention requires 4+1 channels"
        N, _, S, _ = noisy_latents.shape
        D = self.n_pts_per_ray
        dev = noisy_latents.device
        if depth_eps is None:
            depth_eps = torch.randn(N, D, S, S, device=dev)
        override = overwrite_attn_depth is not None
        hw = S * S

        def make(plan, b):
            o = b.ops
            z32 = lambda *s: o.zeros(s, torch.float32)
            i = plan.inputs
            i["noisy"], i["input"], i["eps"] = z32(N, 5, hw), z32(1, 5, hw), z32(N, D, hw)
            i["scal"], i["cams"], i["mask"], i["c"] = z32(2), z32(N + 1, 16), z32(N), z32(1, E.Z_CH)
            i["override"] = z32(N, hw) if override else None
            out = o.empty((N * hw * D, self.output_dim), torch.float32)
            half = 1.0 / float(S)
            E.emit_gridattn(b, noisy=i["noisy"], input_latent=i["input"], depth_override=i["override"], depth_eps=i["eps"],
                            scal=i["scal"], cams=i["cams"], mask=i["mask"], c_embed=i["c"], n_views=N, S=S, D=D, q_first=0,
                            q_count=N, num_layers=self.num_layers, num_heads=self.num_heads, depth_scale=self.depth_scale,
                            depth_shift=self.depth_shift, frustum_out=out,
                            harm_freqs=((2.0 ** torch.arange(7, dtype=torch.float32)) * 0.1).to(o.device),
                            ndc_grid=torch.linspace(1.0 - half, -1.0 + half, S, dtype=torch.float32).to(o.device))
            plan.outputs["y"] = out
            if not override:
                del i["override"]

        plan = self._plan(("fwd", N, S, D, override), make)
        t0 = t.reshape(-1)[0]
        sac = scheduler.sqrt_alphas_cumprod[t0]
        std = scheduler.sqrt_one_minus_alphas_cumprod[t0] / sac / 10.0
        bc, ic = batch_cameras, input_cameras
        cams = pack_cameras(torch.cat([bc.R, ic.R[:1]]), torch.cat([bc.T, ic.T[:1]]),
                            torch.cat([bc.focal_length, ic.focal_length[:1]]),
                            torch.cat([bc.principal_point, ic.principal_point[:1]]))
        feeds = {"noisy": noisy_latents, "input": input_latents[:1], "eps": depth_eps, "scal": torch.stack([sac, std]).float(),
                 "cams": cams, "mask": predict_mask.float(), "c": t_embed[:1]}
        if override:
            feeds["override"] = overwrite_attn_depth.reshape(N, hw)
        return self._execute(plan, feeds).reshape(N, S, S, D, self.output_dim)
