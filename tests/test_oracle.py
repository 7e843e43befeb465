"""CPU: the oracle (oracle/mvd_oracle.py) against the golden outputs of the reference's own modules
(tests/golden/reference_outputs.pt, written by tests/golden/make_golden.py where /root/reference exists), plus the
geometric self-consistency properties that pin the restated pytorch3d arithmetic."""
import os

import pytest
import torch

from common import build_model, rel_l2, state_dict_cpu, synthetic, unet_cfg_of
from oracle import mvd_oracle as O

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_outputs.pt"))
TOL = 2e-5


@pytest.fixture(scope="module")
def small():
    m = build_model(64, 8, D=1, S=32)
    return m, state_dict_cpu(m)


def test_schedule_tables_bit_identical():
    tabs = O.ddpm_tables(1000)
    for steps in (50, 10):
        t = O.ddim_tables(tabs["alphas_cumprod"], steps, 1.0)
        g = GOLD[f"ddim{steps}"]
        for k in ("timesteps", "alphas", "alphas_prev", "sigmas"):
            assert torch.equal(t[k], g[k]), (steps, k)
    t = O.ddim_tables(tabs["alphas_cumprod"], 50, 1.0)
    assert int(t["timesteps"][0]) == 1 and int(t["timesteps"][-1]) == 981
    assert abs(float(t["sigmas"][0]) - 0.0206) < 1e-3 and abs(float(t["sigmas"][-1]) - 0.4545) < 1e-3


def test_unet_small(small):
    m, sd = small
    gold = GOLD["unet"]
    g = torch.Generator().manual_seed(gold["seed"])
    xin = torch.randn(2, 10, 32, 32, generator=g)
    ctx = torch.randn(2, 1, 768, generator=g)
    vol = torch.randn(2, 32, 32, 1, 768, generator=g)
    o = O.unet_forward(sd, xin, torch.tensor([gold["t"]]), ctx, O.volume_pyramid(vol), model_channels=64, num_heads=8,
                       image_size=32, prefix="unet_model.unet_model.")
    assert gold["out"].abs().max() > 0.1  # not the all-zero output of a zero-initialised model
    assert rel_l2(o, gold["out"]) < TOL


@pytest.mark.parametrize("D", [1, 3])
def test_gridattn(D):
    m = build_model(64, 8, D=D, S=32)
    sd = state_dict_cpu(m)
    gold = GOLD[f"gridattn_D{D}"]
    N = gold["N"]
    sc = synthetic.scene_inputs(N, 32, seed=gold["scene_seed"])
    de, _ = synthetic.step_noises(N, D, 32, 1, seed=gold["noise_seed"])
    t = torch.full((N,), gold["t"], dtype=torch.long)
    t_embed = torch.randn(N, 256, generator=torch.Generator().manual_seed(5))
    tables = {k: sd["scheduler." + k] for k in ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod")}
    o = O.gridattn_forward(sd, sc["x_T"] * gold["x_scale"], sc["cams"], torch.ones(N), t_embed, t, tables, de[0],
                           sc["input_latents"], sc["in_cams"], D=D, prefix="view_attn.")
    assert rel_l2(o[:, ::4, ::4, :, ::16], gold["out_sub"]) < TOL
    assert abs(float(o.norm()) / float(gold["out_norm"]) - 1) < TOL


@pytest.mark.parametrize("key,cfg", [("apply_cfg2.5", 2.5), ("apply_cfg1.0", 1.0), ("apply_drop", 1.0)])
def test_apply_model(small, key, cfg):
    m, sd = small
    gold = GOLD[key]
    sc = synthetic.scene_inputs(2, 32, seed=0)
    de, _ = synthetic.step_noises(2, 1, 32, 4, seed=1)
    t = torch.full((2,), gold["t"], dtype=torch.long)
    o = O.apply_model(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], t, de[0],
                      unet_cfg=unet_cfg_of(m), D=1, cfg_scale=cfg, drop_random=gold.get("drop_random"))
    assert rel_l2(o, gold["eps"]) < TOL


def test_ddim_loop(small):
    m, sd = small
    gold = GOLD["ddim4"]
    sc = synthetic.scene_inputs(2, 32, seed=0)
    de, dn = synthetic.step_noises(2, 1, 32, 4, seed=1)
    x, inter = O.ddim_sample(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], de, dn,
                             unet_cfg=unet_cfg_of(m), D=1, num_steps=gold["steps"], eta=1.0, cfg_scale=gold["cfg"],
                             return_intermediates=True)
    for a, b in zip(inter, gold["xt"]):
        assert rel_l2(a["xt"], b) < 1e-4
    assert rel_l2(x, gold["x0"]) < 1e-4


# ---------------------------------------------------------------- geometry properties (pin the pytorch3d restatement)
def test_ray_geometry_self_consistency():
    N, S = 5, 32
    sc = synthetic.scene_inputs(N, S)
    cams = sc["cams"]
    o, d = O.ray_bundle(cams, S)
    C = O.cam_center(cams)
    assert (o - C[:, None, None, :]).abs().max() < 1e-5            # ray origin == camera centre
    d_cam = d.reshape(N, -1, 3) @ cams["R"]                         # direction in the camera frame
    assert (d_cam[..., 2] - 1).abs().max() < 1e-5                   # camera-z component == 1
    z = torch.rand(N, S, S, 1) * 2 + 0.5
    X = o + z * d
    ndc = O.cam_project_ndc(cams, X.reshape(N, -1, 3)[0:1].expand(1, -1, -1).reshape(1, -1, 3))  # view 0's points in every camera
    half = 1.0 / S
    lin = torch.linspace(1 - half, -1 + half, S)
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    own = O.cam_project_ndc({k: v[:1] for k, v in cams.items()}, X[0].reshape(1, -1, 3))
    assert (own[0, :, 0] - gx.reshape(-1)).abs().max() < 2e-5       # own-view reprojection == ray grid
    assert (own[0, :, 1] - gy.reshape(-1)).abs().max() < 2e-5
    assert (1.0 / own[0, :, 2] - z[0].reshape(-1)).abs().max() < 2e-5  # 1 / z_ndc == sampled depth
    inside = ((ndc[..., 0].abs() <= 1) & (ndc[..., 1].abs() <= 1)).float().mean()
    assert 0.4 < float(inside) <= 1.0                               # most cross-view reprojections land in the image


def test_rig_is_rigid_and_relative():
    R, T, f, p = synthetic.gso_rig(8)
    assert torch.allclose(torch.linalg.det(R), torch.ones(9), atol=1e-5)
    assert torch.allclose(R[0], torch.eye(3), atol=1e-6)            # input view has R = I after _get_relative_camera
    C = -torch.einsum("bj,bij->bi", T, R)
    assert torch.allclose(C.norm(dim=-1), torch.full((9,), 1.5), atol=1e-5)


def test_degenerate_attention_identities(small):
    """softmax over a single key == 1: the CLIP cross-attention reduces to to_out(to_v(ctx))."""
    m, sd = small
    p = "unet_model.unet_model.input_blocks.1.1.transformer_blocks.0.attn2"
    x = torch.randn(2, 1024, 64)
    ctx = torch.randn(2, 1, 768)
    full = O.cross_attention(sd, p, x, ctx, 8)
    vec = torch.nn.functional.linear(torch.nn.functional.linear(ctx, sd[p + ".to_v.weight"]), sd[p + ".to_out.0.weight"],
                                     sd[p + ".to_out.0.bias"])
    assert torch.allclose(full, vec.expand(-1, 1024, -1), atol=1e-5)
