"""Pins oracle/mvd_oracle.py against the reference's OWN modules and writes the golden fixtures.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
It imports the unmodified reference (`mvdfusion.*`, `external.sd1.*`, `utils.*`) on CPU through oracle/ref_shims
(stand-ins for the un-installed third-party packages only), loads the product's seeded state dict into the reference
modules with strict=True (which also proves that parameter names and shapes are identical), runs both on the same
seeded inputs and injected noise, asserts oracle == reference to fp32 round-off, and stores the REFERENCE outputs under
tests/golden/*.pt.  tests/test_oracle.py re-checks the oracle against these files everywhere (no reference needed).

Harness-only patches (no reference arithmetic is touched): ViewFusion._init_clip -> no-op (CLIP weights are not
available offline), and torch.normal / torch.randn / torch.randn_like are replaced during a run by functions that pop
pre-drawn tensors, so that the reference consumes exactly the noise the oracle is given.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MVD_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shims"), REF, ROOT, os.path.join(ROOT, "tests")]

from common import build_model, model_config, rel_l2, state_dict_cpu, unet_cfg_of, unet_params  # noqa: E402
from mvdfusion_b200 import synthetic  # noqa: E402
from oracle import mvd_oracle as O  # noqa: E402

import mvdfusion.viewfusion_zero_depth_rgb as ref_vf  # noqa: E402  (the reference)
from mvdfusion.unet import UNetModel as RefUNet  # noqa: E402
from mvdfusion.view_attn_efficient2 import GridAttn as RefGridAttn  # noqa: E402
from pytorch3d.renderer import PerspectiveCameras as ShimCameras  # noqa: E402

TOL = 2e-5


class NoiseQueue:
    """Context manager feeding pre-drawn tensors to the reference's torch.normal / randn / randn_like calls."""

    def __init__(self, normals=(), randns=(), randn_likes=()):
        self.q = {"normal": list(normals), "randn": list(randns), "randn_like": list(randn_likes)}

    def __enter__(self):
        self.saved = (torch.normal, torch.randn, torch.randn_like)
        q = self.q
        torch.normal = lambda mean, std=None, **kw: mean + std * q["normal"].pop(0).reshape(mean.shape)
        torch.randn = lambda *a, **kw: q["randn"].pop(0)
        torch.randn_like = lambda x, **kw: q["randn_like"].pop(0).reshape(x.shape)
        return self

    def __exit__(self, *exc):
        torch.normal, torch.randn, torch.randn_like = self.saved


def shim_cams(c):
    return ShimCameras(R=c["R"], T=c["T"], focal_length=c["f"], principal_point=c["p"])


def check(name, ours, ref, tol=TOL):
    r = rel_l2(ours, ref)
    print(f"  {name:38s} oracle vs reference rel-L2 = {r:.3e}")
    assert r < tol, (name, r)


def build_reference_viewfusion(product_model, mc, heads, D, S):
    cfg = model_config(mc, heads, D, S)["params"]
    cfg["vae_config"] = {"target": "torch.nn.Identity"}
    cfg.pop("ddim_num_steps"), cfg.pop("latent_size")
    ref_vf.ViewFusion._init_clip = lambda self, clip_path: None
    m = ref_vf.ViewFusion(**cfg)
    missing, unexpected = m.load_state_dict(product_model.state_dict(), strict=True)
    assert not missing and not unexpected
    return m.eval()


@torch.no_grad()
def main():
    torch.manual_seed(0)
    out = {}
    mc, heads, S = 64, 8, 32

    # ---------------------------------------------------------------- schedules (full size, cheap)
    print("schedules")
    prod = build_model(mc, heads, D=1, S=S)
    ref = build_reference_viewfusion(prod, mc, heads, 1, S)
    tabs = O.ddpm_tables(1000)
    for k in ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
        assert torch.equal(tabs[k], getattr(ref.scheduler, k)), k
    for steps in (50, 10):
        ref.ddim._make_schedule(steps, "uniform", 1.0, verbose=False)
        t = O.ddim_tables(tabs["alphas_cumprod"], steps, 1.0)
        assert torch.equal(t["timesteps"], torch.from_numpy(ref.ddim.ddim_timesteps.astype("int64")))
        for a, b in (("alphas", ref.ddim.ddim_alphas), ("alphas_prev", ref.ddim.ddim_alphas_prev),
                     ("sigmas", ref.ddim.ddim_sigmas), ("sqrt_one_minus_alphas", ref.ddim.ddim_sqrt_one_minus_alphas)):
            assert torch.equal(t[a], b), (steps, a)
        out[f"ddim{steps}"] = {"timesteps": t["timesteps"], "alphas": t["alphas"], "alphas_prev": t["alphas_prev"],
                               "sigmas": t["sigmas"]}
    print("  DDPM / DDIM tables bit-identical (50 and 10 steps)")

    # ---------------------------------------------------------------- GridAttn, D = 1 and 3
    for D in (1, 3):
        print(f"GridAttn N=3 D={D}")
        N = 3
        prod = build_model(mc, heads, D=D, S=S)
        sd = state_dict_cpu(prod)
        rg = RefGridAttn(in_channels=5, input_size=S, output_dim=768, num_layers=3, z_near_far_scale=0.8, n_pts_per_ray=D)
        rg.load_state_dict({k[len("view_attn."):]: v for k, v in sd.items() if k.startswith("view_attn.")}, strict=True)
        sc = synthetic.scene_inputs(N, S, seed=10 + D)
        de, _ = synthetic.step_noises(N, D, S, 1, seed=20 + D)
        t = torch.full((N,), 621, dtype=torch.long)
        t_embed = torch.randn(N, 256, generator=torch.Generator().manual_seed(5))
        x = sc["x_T"] * 0.7
        with NoiseQueue(normals=[de[0]]):
            r = rg(x, shim_cams(sc["cams"]), torch.ones(N), t_embed, t, ref.scheduler, input_latents=sc["input_latents"],
                   input_cameras=shim_cams(sc["in_cams"]))
        tables = {k: sd["scheduler." + k] for k in ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod")}
        o = O.gridattn_forward(sd, x, sc["cams"], torch.ones(N), t_embed, t, tables, de[0], sc["input_latents"], sc["in_cams"],
                               D=D, prefix="view_attn.")
        check("frustum features", o, r)
        out[f"gridattn_D{D}"] = {"N": N, "t": 621, "scene_seed": 10 + D, "noise_seed": 20 + D, "x_scale": 0.7,
                                 "out_sub": r[:, ::4, ::4, :, ::16].clone(), "out_norm": r.norm()}

    # ---------------------------------------------------------------- UNet / apply_model / DDIM loop (small UNet)
    N, D = 2, 1
    prod = build_model(mc, heads, D=D, S=S)
    sd = state_dict_cpu(prod)
    ref = build_reference_viewfusion(prod, mc, heads, D, S)
    ucfg = unet_cfg_of(prod)
    sc = synthetic.scene_inputs(N, S, seed=0)
    print("UNetModel.forward N=2")
    g = torch.Generator().manual_seed(7)
    xin = torch.randn(N, 10, S, S, generator=g)
    ctx = torch.randn(N, 1, 768, generator=g)
    vol = torch.randn(N, S, S, D, 768, generator=g)
    pyr = ref.unet_model.get_volume_feats_pyramid(vol)
    tt = torch.tensor([301])
    r = ref.unet_model.unet_model(xin, tt, ctx, volume_feats=pyr)
    o = O.unet_forward(sd, xin, tt, ctx, O.volume_pyramid(vol), model_channels=mc, num_heads=heads, image_size=S,
                       prefix="unet_model.unet_model.")
    check("eps (UNetModel)", o, r)
    out["unet"] = {"seed": 7, "t": 301, "out": r.clone()}

    for cfg_scale in (2.5, 1.0):
        print(f"ViewFusion.apply_model N=2 cfg={cfg_scale}")
        de, dn = synthetic.step_noises(N, D, S, 4, seed=1)
        t = torch.full((N,), 501, dtype=torch.long)
        with NoiseQueue(normals=[de[0]]):
            r = ref.apply_model(sc["x_T"], shim_cams(sc["cams"]), sc["input_latents"], shim_cams(sc["in_cams"]),
                                sc["clip_v_embed"], t, cfg_scale=cfg_scale)
        o = O.apply_model(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], t, de[0],
                          unet_cfg=ucfg, D=D, cfg_scale=cfg_scale)
        check("eps (apply_model)", o, r)
        out[f"apply_cfg{cfg_scale}"] = {"t": 501, "eps": r.clone()}

    print("cfg=1.0 with drop_conditions (reference quirk: is_train=True at inference)")
    ref.unet_model.drop_conditions = True
    rnd = torch.tensor([0.03, 0.12])  # view 0: drop all, view 1: drop volume
    saved_rand = torch.rand
    torch.rand = lambda *a, **kw: rnd.clone()
    try:
        with NoiseQueue(normals=[de[0]]):
            r = ref.apply_model(sc["x_T"], shim_cams(sc["cams"]), sc["input_latents"], shim_cams(sc["in_cams"]),
                                sc["clip_v_embed"], t, cfg_scale=1.0)
    finally:
        torch.rand = saved_rand
        ref.unet_model.drop_conditions = False
    o = O.apply_model(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], t, de[0],
                      unet_cfg=ucfg, D=D, cfg_scale=1.0, drop_random=rnd)
    check("eps (apply_model, dropped)", o, r)
    out["apply_drop"] = {"t": 501, "drop_random": rnd, "eps": r.clone()}

    print("DDIMSampler.sample N=2, 4 steps, cfg 2.5")
    steps = 4
    ref.ddim._make_schedule(steps, "uniform", 1.0, verbose=False)
    normals = [de[i] for i in range(steps)]
    likes = [dn[i] for i in range(steps - 1)]
    with NoiseQueue(normals=normals, randns=[sc["x_T"].clone()], randn_likes=likes):
        r, rinter = ref.ddim.sample(shim_cams(sc["cams"]), sc["input_latents"], shim_cams(sc["in_cams"]), sc["clip_v_embed"],
                                    unconditional_scale=2.5, depth=True, return_intermediates=True, verbose=False)
    o, ointer = O.ddim_sample(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], de, dn,
                              unet_cfg=ucfg, D=D, num_steps=steps, eta=1.0, cfg_scale=2.5, return_intermediates=True)
    for a, b in zip(ointer, rinter):
        assert a["t"] == int(b["t"])
        check(f"x_t after t={a['t']}", a["xt"], b["xt"], 1e-4)
    check("x_0 (4-step loop)", o, r, 1e-4)
    out["ddim4"] = {"steps": steps, "cfg": 2.5, "x0": r.clone(), "xt": [b["xt"].clone() for b in rinter]}

    # ---------------------------------------------------------------- full-size UNet, one pass (not stored: 4 GB of weights)
    if os.environ.get("MVD_GOLDEN_FULL", "1") == "1":
        print("full-size UNetModel (320 ch, 1033.8 M params) N=1, oracle vs reference")
        from mvdfusion_b200.mvdfusion.unet import UNetModel
        pu = synthetic.randomize_parameters(UNetModel(**unet_params()), 1234).eval()
        ru = RefUNet(**unet_params())
        ru.load_state_dict(pu.state_dict(), strict=True)
        n_par = sum(p.numel() for p in ru.parameters())
        assert abs(n_par / 1e6 - 1033.79) < 0.01, n_par
        g = torch.Generator().manual_seed(9)
        xin = torch.randn(1, 10, S, S, generator=g)
        ctx = torch.randn(1, 1, 768, generator=g)
        vol = torch.randn(1, S, S, 1, 768, generator=g)
        r = ru.eval()(xin, torch.tensor([981]), ctx, volume_feats=O.volume_pyramid(vol))
        o = O.unet_forward(state_dict_cpu(pu), xin, torch.tensor([981]), ctx, O.volume_pyramid(vol), model_channels=320,
                           num_heads=8, image_size=S)
        check("eps (full-size UNet)", o, r)
        out["unet_full"] = {"seed": 9, "t": 981, "out": r.clone(), "n_params": n_par}

    path = os.path.join(HERE, "reference_outputs.pt")
    torch.save(out, path)
    print("wrote", path, f"({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
