"""Deterministic synthetic inputs and random-init weights for the denoising hot path.

There is no network for datasets or checkpoints, so benchmarks and parity tests use: the GSO camera rig of the reference
(dataset/gso_test.py:116-149: views evenly spaced in azimuth at 30 deg elevation, distance 1.5, focal 2.1875 NDC) made
relative to the input view; seeded Gaussian latents / CLIP embedding / noises; and a seeded re-randomisation of EVERY
parameter — the reference zero-initialises all residual-branch output layers, so a freshly constructed model returns
exactly 0 and any parity check on it would pass vacuously (SURVEY.md §8c "vacuity trap").
"""
import math

import torch


def gso_rig(n_views, elevation_deg=30.0, distance=1.5, focal=2.1875):
    """n_views+1 look-at cameras (index 0 = input view) in pytorch3d conventions: X_cam = X_world @ R + T.
    Restates pytorch3d.look_at_view_transform(dist, elev, azim) with up = +y: camera centre
    C = d (cos e sin a, sin e, cos e cos a), z = normalize(-C), x = normalize(up x z), y = z x x, R = [x y z] (columns), T = -R^T C."""
    n = n_views + 1
    e = math.radians(elevation_deg)
    az = torch.arange(n, dtype=torch.float64) * (2 * math.pi / n) + math.pi / 2
    C = distance * torch.stack([math.cos(e) * torch.sin(az), torch.full_like(az, math.sin(e)), math.cos(e) * torch.cos(az)], -1)
    z = torch.nn.functional.normalize(-C, dim=-1)
    up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64).expand(n, 3)
    x = torch.nn.functional.normalize(torch.cross(up, z, dim=-1), dim=-1)
    y = torch.nn.functional.normalize(torch.cross(z, x, dim=-1), dim=-1)
    R = torch.stack([x, y, z], dim=-1)          # columns
    T = -torch.einsum("bji,bj->bi", R, C)       # -R^T C
    R = torch.einsum("ji,bjk->bik", R[0], R)    # relative to the input view: R' = R_0^T R, T unchanged
    f = torch.full((n, 2), focal, dtype=torch.float64)
    p = torch.zeros(n, 2, dtype=torch.float64)
    return R.float(), T.float(), f.float(), p.float()


def scene_inputs(n_views, S=32, seed=0):
    """x_T (N,5,S,S), input_latents (1,5,S,S) (depth channel zero, viewfusion…py:215), clip_v_embed (N,1,796), cameras."""
    g = torch.Generator().manual_seed(seed)
    R, T, f, p = gso_rig(n_views)
    x_T = torch.randn(n_views, 5, S, S, generator=g)
    inp = torch.cat([torch.randn(1, 4, S, S, generator=g) * 0.8, torch.zeros(1, 1, S, S)], 1)
    clip = torch.randn(1, 1, 768, generator=g).expand(n_views, -1, -1)
    in_e = torch.cat([R[:1].reshape(1, 1, 9), T[:1].reshape(1, 1, 3), f[:1].reshape(1, 1, 2)], -1).expand(n_views, -1, -1)
    b_e = torch.cat([R[1:].reshape(-1, 1, 9), T[1:].reshape(-1, 1, 3), f[1:].reshape(-1, 1, 2)], -1)
    clip_v = torch.cat([clip, in_e, b_e], -1).contiguous()
    cams = {"R": R[1:].contiguous(), "T": T[1:].contiguous(), "f": f[1:].contiguous(), "p": p[1:].contiguous()}
    in_cams = {"R": R[:1].contiguous(), "T": T[:1].contiguous(), "f": f[:1].contiguous(), "p": p[:1].contiguous()}
    return {"x_T": x_T, "input_latents": inp, "clip_v_embed": clip_v, "cams": cams, "in_cams": in_cams}


def step_noises(n_views, D, S, steps, seed=1):
    """Pre-drawn per-step noises: depth jitter (steps,N,D,S,S) and DDIM noise (steps,N,5,S,S)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(steps, n_views, D, S, S, generator=g), torch.randn(steps, n_views, 5, S, S, generator=g)


@torch.no_grad()
def randomize_parameters(module, seed=1234):
    """Seeded re-initialisation of every parameter (no layer left at its zero init): matrices / conv kernels ~ U(-b, b) with
    b = 1/sqrt(fan_in) (PyTorch's default scale), norm gains ~ 1 + 0.1 N(0,1), biases ~ 0.02 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    for name, p in sorted(module.named_parameters(), key=lambda kv: kv[0]):
        if p.dim() >= 2:
            fan_in = p[0].numel()
            b = 1.0 / math.sqrt(fan_in)
            v = (torch.rand(p.shape, generator=g) * 2 - 1) * b
        elif name.endswith("weight"):
            v = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
        else:
            v = 0.02 * torch.randn(p.shape, generator=g)
        p.copy_(v.to(p.device, p.dtype))
    return module
