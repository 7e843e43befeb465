"""CPU fp32 restatement of the CLIP ViT image embedding that precedes the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/mvd_oracle.py's header for who may import oracle/).

Reference call site: `FrozenCLIPImageEmbedder` (external/sd1/ldm/modules/encoders/modules.py:402-441 of the reference):
preprocess (:425-433: bicubic resize to 224 with align_corners=True and no antialias — kornia.geometry.resize is
F.interpolate —, (x + 1) / 2, CLIP mean / std) -> `self.model.encode_image` -> `.float()`; `encode` adds the token axis.
ViewFusion feeds it the input IMAGE in [0, 1] (viewfusion_zero_depth_rgb.py:242), although the embedder documents [-1, 1]: the
(x + 1) / 2 is applied to [0, 1] values — a reference quirk that is reproduced here.

The arithmetic of `encode_image` lives in a third-party dependency that is NOT under /root/reference: OpenAI CLIP
(`clip` from git+https://github.com/openai/CLIP.git, un-pinned in requirements.txt; model 'ViT-L/14').  Restated from its published
clip/model.py: VisionTransformer.forward = conv1 (patch 14, stride 14, no bias) -> [class_embedding ; patches] + positional_embedding
-> ln_pre -> 24 x ResidualAttentionBlock (x + attn(ln_1(x)); x + c_proj(QuickGELU(c_fc(ln_2(x)))), nn.MultiheadAttention with packed
in_proj, QuickGELU(x) = x * sigmoid(1.702 x)) -> ln_post on the class token -> @ proj.  PARITY PIN: tests/test_clip.py checks this
restatement against the independent Hugging Face port (transformers.CLIPVisionModelWithProjection, same published algorithm) with
the weights mapped name by name; the OpenAI package itself is not installable offline, so the pin is cross-implementation,
not against the reference's own dependency — "parity unpinned" at that boundary in the strict sense.
State-dict names are the OpenAI ones under the embedder's prefix: model.visual.{conv1.weight, class_embedding, positional_embedding,
ln_pre.*, transformer.resblocks.N.{ln_1.*, attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.*, ln_2.*, mlp.c_fc.*, mlp.c_proj.*},
ln_post.*, proj}.
"""
import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def preprocess(x, size=224):
    """external/sd1/ldm/modules/encoders/modules.py:425-433"""
    x = F.interpolate(x.float(), size=(size, size), mode="bicubic", align_corners=True, antialias=False)
    x = (x + 1.0) / 2.0
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return (x - mean) / std


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def encode_image(sd, x, heads, prefix="model.visual."):
    """clip/model.py VisionTransformer.forward on a pre-processed batch (B, 3, R, R) -> (B, output_dim)"""
    p = prefix
    w = sd[p + "conv1.weight"]
    patch = w.shape[-1]
    x = F.conv2d(x, w, None, stride=patch)                                   # (B, width, g, g)
    B, C = x.shape[:2]
    x = x.reshape(B, C, -1).permute(0, 2, 1)                                 # (B, g*g, width)
    x = torch.cat([sd[p + "class_embedding"].reshape(1, 1, C).expand(B, -1, -1), x], dim=1) + sd[p + "positional_embedding"]
    x = _ln(sd, p + "ln_pre", x)
    d = C // heads
    i = 0
    while f"{p}transformer.resblocks.{i}.ln_1.weight" in sd:
        q = f"{p}transformer.resblocks.{i}"
        h = _ln(sd, q + ".ln_1", x)
        qkv = F.linear(h, sd[q + ".attn.in_proj_weight"], sd[q + ".attn.in_proj_bias"]).reshape(B, -1, 3, heads, d).permute(2, 0, 3, 1, 4)
        a = ((qkv[0] * d ** -0.5) @ qkv[1].transpose(-1, -2)).softmax(dim=-1) @ qkv[2]
        x = x + F.linear(a.permute(0, 2, 1, 3).reshape(B, -1, C), sd[q + ".attn.out_proj.weight"], sd[q + ".attn.out_proj.bias"])
        h = F.linear(_ln(sd, q + ".ln_2", x), sd[q + ".mlp.c_fc.weight"], sd[q + ".mlp.c_fc.bias"])
        x = x + F.linear(h * torch.sigmoid(1.702 * h), sd[q + ".mlp.c_proj.weight"], sd[q + ".mlp.c_proj.bias"])
        i += 1
    return _ln(sd, p + "ln_post", x[:, 0, :]) @ sd[p + "proj"]


def clip_embed(sd, images, heads, size=224, prefix="model.visual."):
    """FrozenCLIPImageEmbedder.encode (modules.py:435-441): (B, 3, H, W) -> (B, 1, output_dim)"""
    return encode_image(sd, preprocess(images, size), heads, prefix).float().unsqueeze(1)
