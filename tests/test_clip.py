"""CLIP ViT image embedding: FrozenCLIPImageEmbedder (external/sd1/ldm/modules/encoders/modules.py:402-441).

  * oracle/clip_oracle.py (restatement of OpenAI CLIP's published VisionTransformer) is pinned against the independent Hugging Face
    port of the same algorithm, weights mapped name by name (the OpenAI package the reference depends on is not installable offline);
  * the product (patch-embedding / QKV / MLP GEMMs, masked flash attention over the 257-in-272 padded token layout, LayerNorm
    kernels) against the oracle: CPU with the emulated kernels, `-m gpu` with the real ones incl. the full ViT-L/14 size.
"""
import pytest
import torch

from common import rel_l2, synthetic
from oracle import clip_oracle as C

SMALL = dict(width=64, layers=2, heads=4, patch=14, image_size=56, out_dim=48)


def _hf_to_openai(hf_sd, layers):
    """transformers.CLIPVisionModelWithProjection names -> OpenAI clip names under `model.visual.`"""
    v = "vision_model."
    sd = {"model.visual.conv1.weight": hf_sd[v + "embeddings.patch_embedding.weight"],
          "model.visual.class_embedding": hf_sd[v + "embeddings.class_embedding"],
          "model.visual.positional_embedding": hf_sd[v + "embeddings.position_embedding.weight"],
          "model.visual.ln_pre.weight": hf_sd[v + "pre_layrnorm.weight"], "model.visual.ln_pre.bias": hf_sd[v + "pre_layrnorm.bias"],
          "model.visual.ln_post.weight": hf_sd[v + "post_layernorm.weight"], "model.visual.ln_post.bias": hf_sd[v + "post_layernorm.bias"],
          "model.visual.proj": hf_sd["visual_projection.weight"].t().contiguous()}
    for i in range(layers):
        h, o = f"{v}encoder.layers.{i}.", f"model.visual.transformer.resblocks.{i}."
        sd[o + "attn.in_proj_weight"] = torch.cat([hf_sd[h + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
        sd[o + "attn.in_proj_bias"] = torch.cat([hf_sd[h + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0)
        for a, b in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"), ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            sd[o + a + ".weight"], sd[o + a + ".bias"] = hf_sd[h + b + ".weight"], hf_sd[h + b + ".bias"]
    return {k: t.detach().float() for k, t in sd.items()}


def test_clip_oracle_matches_the_huggingface_port():
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(hidden_size=SMALL["width"], intermediate_size=4 * SMALL["width"], num_hidden_layers=SMALL["layers"],
                           num_attention_heads=SMALL["heads"], image_size=SMALL["image_size"], patch_size=SMALL["patch"],
                           projection_dim=SMALL["out_dim"], hidden_act="quick_gelu")
    hf = CLIPVisionModelWithProjection(cfg).eval()
    for p in hf.parameters():  # HF's init leaves some tensors near zero: make every one count
        torch.nn.init.normal_(p, 0.0, 0.3) if p.dim() > 1 else torch.nn.init.normal_(p, 0.5, 0.3)
    sd = _hf_to_openai(hf.state_dict(), SMALL["layers"])
    x = torch.randn(3, 3, SMALL["image_size"], SMALL["image_size"])
    with torch.no_grad():
        want = hf(pixel_values=x).image_embeds
    got = C.encode_image(sd, x, SMALL["heads"])
    assert rel_l2(got, want) < 1e-5


def _product_and_oracle(device, arch, n_img, side):
    from mvdfusion_b200.mvdfusion.clip_encoder import FrozenCLIPImageEmbedder
    m = FrozenCLIPImageEmbedder(**arch)
    synthetic.randomize_parameters(m, 77)
    with torch.no_grad():  # keep the token magnitudes O(1) through the pre-norm stack
        m.model.visual.class_embedding.mul_(0.1)
        m.model.visual.positional_embedding.mul_(0.1)
    m = m.to(device).eval()
    sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
    images = torch.rand(n_img, 3, side, side, generator=torch.Generator().manual_seed(5))
    got = m.encode(images.to(device))
    want = C.clip_embed(sd, images, arch.get("heads", 16), size=arch.get("image_size", 224))
    return got, want


def test_clip_embedder_matches_the_oracle_cpu_emulation(ops_double):
    got, want = _product_and_oracle("cpu", SMALL, 2, 64)
    assert got.shape == want.shape == (2, 1, SMALL["out_dim"])
    assert rel_l2(got, want) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("arch,n_img", [(SMALL, 3), ({}, 1), ({}, 2)])
def test_clip_embedder_matches_the_oracle_gpu(arch, n_img):
    from common import record_parity
    got, want = _product_and_oracle("cuda", arch, n_img, 256)
    assert got.shape == want.shape
    name = "clip_small" if arch else f"clip_vit_l14_b{n_img}"
    assert record_parity(f"{name}_vs_oracle", rel_l2(got, want), 2e-3) < 2e-3


def test_reference_config_target_resolves_to_the_embedder():
    from mvdfusion_b200.config import get_obj_from_str
    cls = get_obj_from_str("external.sd1.ldm.modules.encoders.modules.FrozenCLIPImageEmbedder")
    assert cls.__module__ == "mvdfusion_b200.mvdfusion.clip_encoder"
