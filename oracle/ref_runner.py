"""Runs the reference's OWN implementation of the hot path on the host CPUs, from the sourceless bytecode in oracle/_ref/
(oracle/build_ref.py) with oracle/ref_shims standing in for the un-installed third-party packages.

TEST / BASELINE INFRASTRUCTURE ONLY: used by bench.py (`--impl reference`, and the `cpu_baseline` leg of the native arm) and by
tests/.  Nothing under mvdfusion_b200/ imports it.  Nothing here reads /root/reference.

Harness-only patch (no reference arithmetic is touched): `ViewFusion._init_clip` -> no-op (CLIP weights are not available
offline and the CLIP encoder is outside the timed path: the loop takes the 796-d embedding as an input).
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref", "ref_bytecode.zip")   # sourceless bytecode archive (oracle/build_ref.py), imported by zipimport
SHIMS = os.path.join(HERE, "ref_shims")


def available():
    return os.path.exists(REF_DIR)


def import_reference():
    """Put oracle/ref_shims and oracle/_ref at the head of sys.path and import the reference's facade module."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    for p in (REF_DIR, SHIMS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import mvdfusion.viewfusion_zero_depth_rgb as ref_vf  # the reference (bytecode)
    if not ref_vf.__file__.startswith(REF_DIR):
        raise RuntimeError(f"`mvdfusion` resolved to {ref_vf.__file__}, not to oracle/_ref")
    ref_vf.ViewFusion._init_clip = lambda self, clip_path: None
    return ref_vf


def build_reference_model(model_params, seed=1234):
    """The reference's ViewFusion from the `model.params` of configs/mvd_gso.yaml (as tests/common.model_config writes them),
    with the seeded re-randomisation of every parameter that the product and the oracle use (SURVEY.md §8c vacuity trap)."""
    from mvdfusion_b200 import synthetic
    ref_vf = import_reference()
    cfg = dict(model_params)
    cfg["vae_config"] = {"target": "torch.nn.Identity"}
    steps = cfg.pop("ddim_num_steps", 50)
    S = cfg.pop("latent_size", 32)
    m = ref_vf.ViewFusion(**cfg)
    if S != 32:  # the reference hard-codes latent_size = 32 (viewfusion_zero_depth_rgb.py:92); SURVEY.md §5 lists the three knobs
        m.ddim.latent_size = S
    m.ddim._make_schedule(steps, "uniform", 1.0, verbose=False)
    synthetic.randomize_parameters(m, seed)
    return m.eval()


def shim_cameras(c):
    from pytorch3d.renderer import PerspectiveCameras  # oracle/ref_shims
    return PerspectiveCameras(R=c["R"], T=c["T"], focal_length=c["f"], principal_point=c["p"])


@torch.no_grad()
def time_denoising_steps(model, scene, cfg_scale, steps, warmup, budget_s=240.0, num_ddim=50):
    """The body of DDIMSampler.sample's loop (mvdfusion/sampler.py:119-142: time_steps -> denoise_apply -> apply_model ->
    denoise_apply_impl), iteration i at DDIM index num_ddim-1-i, on the reference's own modules.  Runs `warmup` untimed and up
    to `steps` timed iterations, stopping early once `budget_s` of wall time is spent.  Returns (list of seconds, x)."""
    cams, in_cams = shim_cameras(scene["cams"]), shim_cameras(scene["in_cams"])
    x = scene["x_T"].clone()
    B = x.shape[0]
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        index = num_ddim - 1 - (i % num_ddim)
        step = int(model.ddim.ddim_timesteps[index])
        t0 = time.perf_counter()
        ts = torch.full((B,), step, dtype=torch.long)
        x, _ = model.ddim.denoise_apply(x, cams, scene["input_latents"], in_cams, scene["clip_v_embed"], ts, index,
                                        is_step0=index == 0, cfg_scale=cfg_scale)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if i >= warmup and time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    return times, x
