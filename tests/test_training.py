"""Training forward + backward: ViewFusion.forward -> p_losses -> mvdfusion_b200.training, whose contractions
(forward, dgrad, wgrad) run on the library's GEMM kernel behind autograd Functions.  Oracle: torch.autograd through the fp32 CPU
restatement (oracle/mvd_oracle.py) on identical inputs and injected draws.  Gate: rel-L2 <= 1e-2 on every parameter gradient group
(fp16 operand rounding in three GEMMs per layer), loss value <= 1e-3.

CPU (`not gpu`): the kernels are emulated by tests/ops_double.py (host logic, operand packing, role swaps, im2col^T, flipped kernels).
`-m gpu`: the same comparison with the real sm_100a kernels, plus the native Function primitives against torch.autograd."""
import os

import pytest
import torch
import torch.nn.functional as F

from common import build_model, rel_l2, state_dict_cpu, synthetic, synthetic_dataset_batch, standin_clip_encode, standin_vae_encode, unet_cfg_of
from oracle import mvd_oracle as O

TRAINER = {"input_batch_size": 1, "train_batch_size": 2, "random_views": False}


def _scene_batch(m, side=256):
    batch = synthetic_dataset_batch(5, side, seed=11)
    images = batch.pop("images")
    batch["latents"] = standin_vae_encode(images, m.z_scale_factor) * 4.0     # 32x32 latents of unit-ish scale
    batch["clip_embed"] = standin_clip_encode(images)
    return batch


def _oracle_loss_and_grads(m, batch, t, noise, depth_eps, D):
    """the reference algorithm (oracle, fp32 CPU) under torch.autograd with the product's parameters as leaves"""
    sd = {k: v.detach().clone().float().cpu() for k, v in m.state_dict().items()}
    names = [k for k, p in m.named_parameters() if p.requires_grad]
    for k in names:
        sd[k].requires_grad_(True)
    mc = build_model(64, 8, D=D, S=32)                                        # CPU twin only for prepare_batch's host logic
    bl, bc, il, ic, cv = mc.prepare_batch({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in batch.items()}, TRAINER)
    tc = t.cpu()
    noisy = mc.scheduler.sqrt_alphas_cumprod[tc].view(-1, 1, 1, 1) * bl + mc.scheduler.sqrt_one_minus_alphas_cumprod[tc].view(-1, 1, 1, 1) * noise.cpu()
    cams = {"R": bc.R, "T": bc.T, "f": bc.focal_length, "p": bc.principal_point}
    icams = {"R": ic.R, "T": ic.T, "f": ic.focal_length, "p": ic.principal_point}
    pred = O.apply_model(sd, noisy, cams, il, icams, cv, tc, depth_eps.cpu(), unet_cfg=unet_cfg_of(m), D=D, cfg_scale=1.0)
    loss = F.mse_loss(noise.cpu(), pred)
    grads = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
    return float(loss), dict(zip(names, grads))


def _group(name):
    if name.startswith("view_attn."):
        return "view_attn"
    if ".aligned_attn_" in name:
        return "unet.aligned_attn"
    if name.startswith("unet_model."):
        return "unet.attn" if any(s in name for s in (".transformer_blocks.", ".proj_in.", ".proj_out.", ".norm.")) else "unet.other"
    return name.split(".")[0]


def _check(m, dev, D):
    torch.manual_seed(0)
    for p in m.parameters():
        p.requires_grad_(True)
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in _scene_batch(m).items()}
    g = torch.Generator().manual_seed(3)
    t = torch.full((2,), 431, dtype=torch.long, device=dev)
    noise = torch.randn(2, 5, 32, 32, generator=g).to(dev)
    depth_eps = torch.randn(2, D, 32, 32, generator=g).to(dev)
    m.train()
    loss = m.forward(batch, TRAINER) if False else m.p_losses(batch, TRAINER, t=t, noise=noise, depth_eps=depth_eps)
    assert loss.requires_grad
    loss.backward()
    ref_loss, ref = _oracle_loss_and_grads(m, batch, t, noise, depth_eps, D)
    assert abs(float(loss) - ref_loss) <= 1e-3 * abs(ref_loss), (float(loss), ref_loss)
    groups = {}
    for k, p in m.named_parameters():
        r = ref.get(k)
        if r is None:                 # parameters the path never reads (view_attn.t_embedder / ray_embedder: SURVEY.md §2.3)
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        a, b = groups.setdefault(_group(k), ([], []))
        a.append(p.grad.detach().float().cpu().reshape(-1))
        b.append(r.reshape(-1))
    out = {}
    for gname, (a, b) in groups.items():
        out[gname] = rel_l2(torch.cat(a), torch.cat(b))
    return float(loss), ref_loss, out


@pytest.mark.parametrize("D", [3])
def test_training_gradients_match_autograd_on_the_oracle_cpu_emulation(ops_double, D):
    m = build_model(64, 8, D=D, S=32)
    loss, ref_loss, rel = _check(m, "cpu", D)
    print("loss", loss, ref_loss, rel)
    assert set(rel) >= {"view_attn", "unet.aligned_attn", "unet.attn", "unet.other", "cc_projection", "time_embed"}
    for gname, r in rel.items():
        assert r < 1e-2, (gname, r)


@pytest.mark.gpu
@pytest.mark.parametrize("D", [3, 1])
def test_training_gradients_match_autograd_on_the_oracle_gpu(D):
    from common import record_parity
    m = build_model(64, 8, D=D, S=32, device="cuda")
    loss, ref_loss, rel = _check(m, "cuda", D)
    for gname, r in rel.items():
        record_parity(f"train_grad_D{D}_{gname}_vs_oracle_autograd", r, 1e-2)
        assert r < 1e-2, (gname, r)


@pytest.mark.gpu
def test_contraction_functions_forward_dgrad_wgrad():
    """the autograd Functions over mvd_gemm_f16 against torch.autograd in fp32 (shapes of the step: linear with a ragged K,
    3x3 convolution with the stem's 10 and the head's 5 channels, stride-2 convolution)"""
    from mvdfusion_b200 import training as T
    g = torch.Generator().manual_seed(1)
    r = lambda *s: torch.randn(*s, generator=g).cuda().requires_grad_(True)
    # linear 723 -> 256 (GridAttn pre_layer_b: K is not a multiple of 8)
    x, w, b = r(640, 723), r(256, 723), r(256)
    y = T.linear(x, w, b)
    ref = F.linear(x, w, b)
    gy = torch.randn_like(ref)
    for got, want in zip(torch.autograd.grad(y, (x, w, b), gy), torch.autograd.grad(ref, (x, w, b), gy)):
        assert rel_l2(got, want) < 2e-3
    assert rel_l2(y, ref) < 2e-3
    for cin, cout in ((10, 64), (64, 5), (128, 192)):
        x, w, b = r(2, 16, 16, cin), r(cout, cin, 3, 3), r(cout)
        y = T.conv3x3(x, w, b)
        ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=1).permute(0, 2, 3, 1)
        gy = torch.randn_like(ref)
        assert rel_l2(y, ref) < 2e-3
        for got, want in zip(torch.autograd.grad(y, (x, w, b), gy), torch.autograd.grad(ref, (x, w, b), gy)):
            assert rel_l2(got, want) < 2e-3, (cin, cout)
    x, w, b = r(2, 16, 16, 64), r(64, 64, 3, 3), r(64)
    y = T.conv3x3_stride2(x, w, b)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, stride=2, padding=1).permute(0, 2, 3, 1)
    gy = torch.randn_like(ref)
    assert rel_l2(y, ref) < 2e-3
    for got, want in zip(torch.autograd.grad(y, (x, w, b), gy), torch.autograd.grad(ref, (x, w, b), gy)):
        assert rel_l2(got, want) < 2e-3


# ------------------------------------------------------------------------------------------------ ABI 15: normalisation / activation kernels
def _pointwise_cases():
    """(name, Function-level callable, torch reference, inputs) at the shapes of the training graph: LayerNorm with / without the affine
    part, GroupNorm32 (+ SiLU) with 2 / 10 / 40 channels per group and a ragged last pixel chunk, GELU / SiLU with a scalar tail, GEGLU"""
    from mvdfusion_b200 import training as T
    g = torch.Generator().manual_seed(5)
    r = lambda *s, scale=1.0, shift=0.0: torch.randn(*s, generator=g) * scale + shift
    cases = []
    for rows, C, affine in ((300, 320, True), (64, 1280, True), (1000, 256, False), (37, 64, True)):
        ins = [r(rows, C, scale=2.0, shift=0.5)] + ([r(C, shift=1.0), r(C)] if affine else [])
        cases.append((f"layernorm_{rows}x{C}_{'affine' if affine else 'plain'}",
                      (lambda x, *gb: T.layer_norm(x, gb[0] if gb else None, gb[1] if gb else None, 1e-5)),
                      (lambda x, *gb: F.layer_norm(x, (x.shape[-1],), gb[0] if gb else None, gb[1] if gb else None, 1e-5)), ins))
    for n, H, W, C, act in ((2, 8, 8, 64, True), (3, 5, 7, 320, True), (2, 16, 16, 1280, False), (1, 4, 4, 2560, True), (2, 32, 32, 320, False),
                            (2, 4, 4, 32, True), (1, 9, 9, 960, True)):
        ins = [r(n, H, W, C, scale=1.5, shift=0.3), r(C, shift=1.0), r(C)]
        ref = lambda x, ga, be, act=act: (lambda y: F.silu(y) if act else y)(F.group_norm(x.permute(0, 3, 1, 2), 32, ga, be, 1e-5).permute(0, 2, 3, 1))
        cases.append((f"groupnorm_{n}x{H}x{W}x{C}_{'silu' if act else 'plain'}", (lambda x, ga, be, act=act: T._GroupNormFn.apply(x, ga, be, 1e-5, act)), ref, ins))
    # bilinear gather (grid_sample, border clamp, align_corners): coordinates beyond [-1, 1] exercise the clamp; xy is not differentiated
    for V, H, W, C, Pn in ((3, 8, 8, 256, 1000), (1, 32, 32, 64, 333), (2, 5, 7, 12, 50)):
        xy = (torch.rand(V, Pn, 2, generator=g) * 2.6 - 1.3)
        xy[:, :4] = torch.tensor([[-1.0, -1.0], [1.0, 1.0], [1.0, -1.0], [0.0, 0.0]])      # exact corners / centre
        ref = lambda m, xy=xy: F.grid_sample(m.permute(0, 3, 1, 2), xy.to(m.dtype).unsqueeze(2), mode="bilinear", padding_mode="border",
                                             align_corners=True)[..., 0].permute(0, 2, 1)
        cases.append((f"gather_{V}x{H}x{W}x{C}_P{Pn}", (lambda m, xy=xy: T.bilinear_gather(m, xy.to(m.device))), ref, [r(V, H, W, C)]))
    cases.append(("gelu_tail", T.gelu, F.gelu, [r(7, 331, scale=2.0)]))
    cases.append(("gelu_vec", T.gelu, F.gelu, [r(64, 512, scale=2.0)]))
    cases.append(("silu", T.silu, F.silu, [r(2, 1280, scale=3.0)]))
    cases.append(("geglu", T.geglu, (lambda h: h.chunk(2, dim=-1)[0] * F.gelu(h.chunk(2, dim=-1)[1])), [r(5, 40, 2560, scale=1.5)]))
    return cases


def _check_pointwise(dev, tol, place=None):
    """every case runs (one failing kernel does not hide the others); returns name -> worst rel-L2 over (y, grads).
    place: how the inputs are put in front of the kernels (the CPU-shim runs pass common.guarded_clone: buffers that end at a guard page)"""
    worst, bad = {}, []
    place = place or (lambda t: t.clone().to(dev))
    for name, fn, ref, ins in _pointwise_cases():
        a = [place(t).requires_grad_(True) for t in ins]
        b = [t.clone().double().requires_grad_(True) for t in ins]
        y, yr = fn(*a), ref(*b)
        gy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(9))
        got = torch.autograd.grad(y, a, gy.to(dev))
        want = torch.autograd.grad(yr, b, gy.double())
        errs = [rel_l2(y.detach().cpu(), yr.detach().float())] + [rel_l2(p.cpu(), q.float()) for p, q in zip(got, want)]
        finite = bool(torch.isfinite(y).all()) and all(bool(torch.isfinite(p).all()) for p in got)
        print(f"pointwise {name}: y / grads rel-L2 {['%.2e' % e for e in errs]} finite={finite}")
        worst[name] = max(errs)
        if not finite or not max(errs) < tol:
            bad.append((name, errs, finite))
    assert not bad, bad
    return worst


def test_pointwise_functions_cpu_emulation_matches_autograd(ops_double):
    """pins the closed forms of tests/ops_double.py (and the Functions' host logic: shapes, saved tensors, optional affine part)
    to torch.autograd in float64"""
    _check_pointwise("cpu", 2e-5)


@pytest.mark.gpu
def test_pointwise_kernels_forward_backward_gpu():
    """mvd_layernorm / groupnorm / act _{fwd,bwd}_f32 (csrc/train.cu) through their autograd Functions against torch.autograd in
    float64: outputs and every gradient (dx, dgamma, dbeta) at rel-L2 <= 2e-5 (fp32 kernels; atomics change the summation order only)"""
    from common import record_parity
    for name, err in _check_pointwise("cuda", 2e-5).items():
        record_parity(f"train_pointwise_{name}_vs_autograd_f64", err, 2e-5)


# ------------------------------------------------------------------------------------------------ the kernel SOURCE on the CPU shim
def test_pointwise_kernel_source_on_the_cpu_shim(tmp_path, monkeypatch):
    """The CUDA kernels of csrc/train.cu themselves — not an emulation of their semantics — executed on host threads
    (tests/native/cpu_emul/cuda_on_cpu.h: real threads per block, barriers for __syncthreads / warp shuffles, serialised atomics),
    bound through the product's own ops.NativeOps methods and autograd Functions, against torch.autograd in float64.  Checks the
    indexing, reductions, ragged chunks and closed forms in a container without a GPU; the -m gpu twin checks the nvcc build."""
    import mvdfusion_b200.runtime as rt
    from common import build_cpu_shim, shim_ops
    shim = shim_ops(build_cpu_shim(["train.cu"], tmp_path), monkeypatch)
    monkeypatch.setattr(rt, "get_ops", lambda dev: shim)
    from common import guarded_clone
    worst = _check_pointwise("cpu", 2e-5, place=guarded_clone)
    assert len(worst) >= 18


@pytest.mark.parametrize("order", ["reverse", "shuffle:3"])
def test_pointwise_kernel_source_is_schedule_independent(tmp_path, monkeypatch, order):
    """racecheck stand-in for csrc/train.cu (no compute-sanitizer run exists for these kernels: the GPU budget went elsewhere): the shim
    resumes threads in another order between barriers; every case still matches float64 autograd"""
    import mvdfusion_b200.runtime as rt
    from common import build_cpu_shim, shim_ops
    shim = shim_ops(build_cpu_shim(["train.cu"], tmp_path), monkeypatch)
    monkeypatch.setattr(rt, "get_ops", lambda dev: shim)
    monkeypatch.setenv("MVD_SHIM_ORDER", order)
    _check_pointwise("cpu", 2e-5)
