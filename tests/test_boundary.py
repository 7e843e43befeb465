"""CPU: the drop-in boundary against the reference's OWN files.

  * configs/mvd_gso.yaml and configs/mvd_train.yaml of the reference load UNCHANGED (only the weight paths, which point at files that
    do not exist offline, are blanked) through mvdfusion_b200.config AND through the reference's own utils/load_model.py after the
    sys.modules recipe of INTEGRATION.md §1 — skipped when /root/reference is absent (the GPU box).
  * ViewFusion.prepare_batch against the golden written by the REFERENCE's prepare_batch (tests/golden/make_golden_prepare_batch.py).
"""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from common import ROOT, fast_init, standin_clip_encode, standin_vae_encode, synthetic_dataset_batch

REF = os.environ.get("MVD_REFERENCE", "/root/reference")
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "configs")), reason="the reference checkout is not on this box")
WEIGHT_KEYS = ("vae_path", "clip_path", "unet_path", "unet_cc_path")
EXPECT = {"mvd_gso.yaml": dict(D=1, finetune_unet=False), "mvd_train.yaml": dict(D=3, finetune_unet=True)}


@needs_reference
@pytest.mark.parametrize("name", sorted(EXPECT))
def test_reference_yaml_loads_through_the_product_config(name):
    from mvdfusion_b200.config import instantiate_from_config, load_yaml
    cfg = load_yaml(os.path.join(REF, "configs", name))
    for k in WEIGHT_KEYS:
        cfg["model"]["params"][k] = None
    with fast_init():  # shapes and names only: the 1.1 B parameters are allocated, not drawn
        m = instantiate_from_config(cfg["model"])
    assert type(m).__module__ == "mvdfusion_b200.mvdfusion.viewfusion_zero_depth_rgb" and type(m).__name__ == "ViewFusion"
    assert m.view_attn.n_pts_per_ray == EXPECT[name]["D"] and m.finetune_unet == EXPECT[name]["finetune_unet"]
    n_unet = sum(p.numel() for p in m.unet_model.unet_model.parameters())
    n_grid = sum(p.numel() for p in m.view_attn.parameters())
    assert abs(n_unet / 1e6 - 1033.79) < 0.01 and abs(n_grid / 1e6 - 3.28) < 0.01, (n_unet, n_grid)
    assert sum(p.numel() for p in m.vae.parameters()) / 1e6 == pytest.approx(83.65, abs=0.01)
    sd = m.state_dict()
    for k in ("view_attn.aggregation_transformer.layer_list.2.attn.qkv.bias", "unet_model.unet_model.out.2.bias",
              "unet_model.unet_model.middle_block.2.aligned_attn_proj_in.weight", "scheduler.alphas_cumprod", "cc_projection.4.weight",
              "vae.decoder.conv_out.weight", "time_embed.2.bias"):
        assert k in sd, k
    # the sections demo.py / train.py read next to `model` stay reachable unchanged
    assert cfg["trainer"]["train_batch_size"] >= 1 and "dataset" in cfg and "saver" in cfg


@needs_reference
def test_reference_load_model_builds_the_product_after_the_integration_recipe():
    """INTEGRATION.md §1 verbatim in a fresh interpreter: alias the mirror under the reference's module names, then call the
    REFERENCE's utils.load_model.instantiate_from_config on the reference's yaml."""
    code = textwrap.dedent(f"""
        import sys
        sys.path[:0] = [{os.path.join(ROOT, 'oracle', 'ref_shims')!r}, {REF!r}, {ROOT!r}]
        import torch, mvdfusion_b200.mvdfusion as _m
        import mvdfusion_b200.mvdfusion.viewfusion_zero_depth_rgb, mvdfusion_b200.mvdfusion.view_attn_efficient2
        import mvdfusion_b200.mvdfusion.unet, mvdfusion_b200.mvdfusion.scheduler, mvdfusion_b200.mvdfusion.sampler
        import mvdfusion_b200.mvdfusion.attention, mvdfusion_b200.mvdfusion.embedder
        sys.modules["mvdfusion"] = _m
        for name in ("viewfusion_zero_depth_rgb", "view_attn_efficient2", "unet", "scheduler", "sampler", "attention", "embedder"):
            sys.modules[f"mvdfusion.{{name}}"] = getattr(_m, name)
        from omegaconf import OmegaConf                       # oracle/ref_shims stand-in: yaml.safe_load
        from utils.load_model import instantiate_from_config  # the reference's own factory
        assert instantiate_from_config.__code__.co_filename.startswith({REF!r}) or 'load_model' in instantiate_from_config.__code__.co_filename
        cfg = OmegaConf.load({os.path.join(REF, 'configs', 'mvd_gso.yaml')!r})
        for k in {WEIGHT_KEYS!r}:
            cfg["model"]["params"][k] = None
        sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
        from common import fast_init
        with fast_init():
            m = instantiate_from_config(cfg["model"])
        print(type(m).__module__, type(m).__name__, sum(p.numel() for p in m.parameters()))
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    mod, cls, n = out.stdout.strip().splitlines()[-1].split()
    assert (mod, cls) == ("mvdfusion_b200.mvdfusion.viewfusion_zero_depth_rgb", "ViewFusion")
    assert abs(int(n) / 1e6 - 1122.65) < 0.5, n


@pytest.mark.parametrize("case", ["linspace", "random"])
def test_prepare_batch_matches_the_reference_golden(case):
    from common import build_model
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "prepare_batch_outputs.pt"))[case]
    m = build_model(64, 8, D=1, S=32)
    batch = synthetic_dataset_batch(gold["n_views"], 64, seed=3)
    images = batch.pop("images")
    batch["latents"] = standin_vae_encode(images, m.z_scale_factor)   # pre-encoded inputs: the frozen encoders are outside this row
    batch["clip_embed"] = standin_clip_encode(images)
    gen = torch.Generator().manual_seed(gold["generator_seed"]) if gold["generator_seed"] is not None else None
    bl, bc, il, ic, cv = m.prepare_batch(batch, gold["trainer_config"], generator=gen)
    for name, got in (("batch_latents", bl), ("input_latents", il), ("clip_v_embed", cv), ("batch_R", bc.R), ("batch_T", bc.T),
                      ("batch_f", bc.focal_length), ("batch_p", bc.principal_point), ("input_R", ic.R), ("input_T", ic.T)):
        want = gold[name]
        assert got.shape == want.shape, name
        assert torch.allclose(got, want, atol=2e-6, rtol=0), (name, float((got - want).abs().max()))
