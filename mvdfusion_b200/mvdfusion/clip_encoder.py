"""external/sd1/ldm/modules/encoders/modules.py:402-441 of the reference: `FrozenCLIPImageEmbedder`, the frozen CLIP ViT-L/14 image
encoder that produces the 768-d embedding of the input view once per scene (viewfusion_zero_depth_rgb.py:103-105,154-155,242) —
SURVEY.md §8f rank 3, the step before `prepare_batch`'s 796-d [clip | cameras] vector.

Parameter names are OpenAI CLIP's (`model.visual.conv1.weight`, `model.visual.transformer.resblocks.N.attn.in_proj_weight`, …) so that
the `clip_image_encoder.*` entries of the reference's checkpoints load unchanged.  `forward` / `encode` run on the library:
patch embedding, QKV / out-proj / MLP GEMMs (tcgen05), flash attention with the 257-token sequence in a 272-row padded layout
(mvd_attn_self_masked_f16), LayerNorm kernels; QuickGELU(x) = silu(1.702 x) / 1.702 is folded into the packed c_fc / c_proj weights.
The image pre-processing (bicubic 224 resize, CLIP normalisation, patch extraction) is torch data movement at the boundary.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import engine as E
from .. import runtime
from ..runtime import PlanCache, WeightCache, current_stream

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class _Params(nn.Module):
    """bare parameter holder (the names are the contract; the arithmetic lives in emit_clip_vit)"""


def _visual(width, layers, patch, grid, out_dim):
    v = _Params()
    v.conv1 = nn.Conv2d(3, width, patch, stride=patch, bias=False)
    scale = width ** -0.5
    v.class_embedding = nn.Parameter(scale * torch.randn(width))
    v.positional_embedding = nn.Parameter(scale * torch.randn(grid * grid + 1, width))
    v.ln_pre = nn.LayerNorm(width)
    v.transformer = _Params()
    blocks = []
    for _ in range(layers):
        b = _Params()
        b.ln_1 = nn.LayerNorm(width)
        b.attn = _Params()
        b.attn.in_proj_weight = nn.Parameter(torch.randn(3 * width, width) * scale)
        b.attn.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        b.attn.out_proj = nn.Linear(width, width)
        b.ln_2 = nn.LayerNorm(width)
        b.mlp = _Params()
        b.mlp.c_fc = nn.Linear(width, 4 * width)
        b.mlp.c_proj = nn.Linear(4 * width, width)
        blocks.append(b)
    v.transformer.resblocks = nn.ModuleList(blocks)
    v.ln_post = nn.LayerNorm(width)
    v.proj = nn.Parameter(scale * torch.randn(width, out_dim))
    return v


class FrozenCLIPImageEmbedder(nn.Module):
    """modules.py:402-441.  model: 'ViT-L/14' (the only one the reference uses) or a path to its checkpoint (OpenAI TorchScript
    archive or a state dict).  width / layers / heads / image size are overridable for small-size parity cases."""

    ARCH = {"ViT-L/14": dict(width=1024, layers=24, heads=16, patch=14, image_size=224, out_dim=768)}

    def __init__(self, model="ViT-L/14", jit=False, device="cpu", antialias=False, **arch):
        super().__init__()
        a = dict(self.ARCH["ViT-L/14"])
        a.update(arch)
        self.width, self.layers, self.heads, self.patch, self.image_size, self.out_dim = (a[k] for k in ("width", "layers", "heads", "patch", "image_size", "out_dim"))
        self.grid = self.image_size // self.patch
        self.antialias = antialias
        self.model = _Params()
        self.model.visual = _visual(self.width, self.layers, self.patch, self.grid, self.out_dim)
        self.register_buffer("mean", torch.tensor(CLIP_MEAN), persistent=False)
        self.register_buffer("std", torch.tensor(CLIP_STD), persistent=False)
        self._cache = WeightCache()
        self._plans = PlanCache(2)
        if isinstance(model, str) and model not in self.ARCH and model:
            self._load(model)

    def _load(self, path):
        try:
            sd = torch.jit.load(path, map_location="cpu").state_dict()
        except RuntimeError:
            sd = torch.load(path, map_location="cpu")
            sd = sd.get("state_dict", sd)
        sd = {("model." + k if k.startswith("visual.") else k): v.float() for k, v in sd.items() if "visual." in k}
        missing, _ = self.load_state_dict(sd, strict=False)
        if missing:
            raise RuntimeError(f"CLIP checkpoint {path} lacks {len(missing)} visual parameters, e.g. {missing[:3]}")

    def preprocess(self, x):
        """modules.py:425-433 (expects [-1, 1]; ViewFusion passes [0, 1] images — reproduced as is)"""
        x = F.interpolate(x.float(), size=(self.image_size, self.image_size), mode="bicubic", align_corners=True, antialias=self.antialias)
        x = (x + 1.0) / 2.0
        return (x - self.mean.view(1, 3, 1, 1)) / self.std.view(1, 3, 1, 1)

    @torch.no_grad()
    def forward(self, x):
        if isinstance(x, list):  # [""] denotes condition dropout for ucg (modules.py:437-440)
            return torch.zeros(1, self.out_dim, device=self.mean.device)
        dev = x.device
        ops = runtime.get_ops(dev)
        B = x.shape[0]
        xp = self.preprocess(x)
        g, P = self.grid, self.patch
        # patch extraction: (B, 3, gP, gP) -> rows (B*g*g, 3*P*P) in conv1's (c, py, px) order, K padded to a multiple of 8
        cols = xp.reshape(B, 3, g, P, g, P).permute(0, 2, 4, 1, 3, 5).reshape(B * g * g, 3 * P * P)
        kp = E._round_up(cols.shape[1], 8)
        key = (B, str(dev))
        sd = self._cache.get(self, ops)
        if key not in self._plans or self._plans[key][0] is not sd:
            self._plans[key] = (sd, E.ClipPlan(ops, E.PackedWeights(sd, ops, "model.visual."), B, self.width, self.layers, self.heads, g * g + 1, kp, self.out_dim))
        plan = self._plans[key][1]
        plan.patches.zero_()
        plan.patches[:, :cols.shape[1]].copy_(cols.half())
        plan.prog.run(current_stream(dev))
        return plan.out.clone().float()

    def encode(self, im):
        return self(im).unsqueeze(1)
