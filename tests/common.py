"""Shared helpers of the test-suite (model configs, building the product model, rel-L2)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mvdfusion_b200 import synthetic  # noqa: E402


def unet_params(model_channels=320, num_heads=8, image_size=32):
    """configs/mvd_gso.yaml:30-46 of the reference (model_channels / heads reducible for small-size parity cases)."""
    return dict(image_size=image_size, in_channels=10, out_channels=5, model_channels=model_channels,
                attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=num_heads,
                use_spatial_transformer=True, use_view_aligned_transformer=True, transformer_depth=1, context_dim=768,
                use_checkpoint=True, legacy=False)


def model_config(model_channels=320, num_heads=8, D=1, S=32, drop_conditions=False, ddim_steps=50):
    """model: section of configs/mvd_gso.yaml (target strings exactly as the reference writes them)."""
    return {
        "target": "mvdfusion.viewfusion_zero_depth_rgb.ViewFusion",
        "params": {
            "vae_path": None, "clip_path": None, "unet_path": None, "z_scale_factor": 0.18215, "objective": "noise",
            "loss_type": "l2", "embed_camera_pose": True, "finetune_projection": True, "finetune_unet": False,
            "finetune_cross_attn": True, "finteune_view_attn": True, "drop_conditions": drop_conditions,
            "ddim_num_steps": ddim_steps, "latent_size": S,
            "view_attn_config": {"target": "mvdfusion.view_attn_efficient2.GridAttn",
                                 "params": {"in_channels": 5, "input_size": S, "output_dim": 768, "num_layers": 3,
                                            "z_near_far_scale": 0.8, "n_pts_per_ray": D}},
            "unet_config": {"target": "mvdfusion.unet.UNetModel", "params": unet_params(model_channels, num_heads, S)},
            "ddpm_config": {"target": "mvdfusion.scheduler.DDPMScheduler", "params": {"timesteps": 1000}},
            "vae_config": None,
        },
    }


def build_model(model_channels=320, num_heads=8, D=1, S=32, seed=1234, device="cpu", **kw):
    from mvdfusion_b200.config import instantiate_from_config
    m = instantiate_from_config(model_config(model_channels, num_heads, D, S, **kw))
    synthetic.randomize_parameters(m, seed)
    return m.to(device).eval()


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def state_dict_cpu(model):
    return {k: v.detach().float().cpu() if v.is_floating_point() else v.detach().cpu() for k, v in model.state_dict().items()}


def unet_cfg_of(model):
    um = model.unet_model.unet_model
    return {"model_channels": um.model_channels, "num_heads": um.num_heads, "image_size": um.image_size,
            "channel_mult": list(um.channel_mult)}


def record_parity(name, value, tol):
    """Append a measured rel-L2 to gpurun_out/parity.jsonl (brought back from the GPU box) and echo it."""
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, "rel_l2": value, "tolerance": tol}) + "\n")
    except OSError:
        pass
    print(f"parity {name}: rel-L2 {value:.3e} (tolerance {tol:g})")
    return value
