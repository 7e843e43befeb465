"""-m gpu: every C-ABI entry point on the B200 against the CPU emulation of its documented semantics
(tests/ops_double.py), on random inputs at the shapes the hot path uses."""
import pytest
import torch

from mvdfusion_b200 import ops as OPS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    return OPS.NativeOps("cuda:0")


@pytest.fixture(scope="module")
def dbl():
    from ops_double import TorchOpsDouble
    return TorchOpsDouble()


def stream_of(nat):
    return torch.cuda.current_stream().cuda_stream if nat.device.type == "cuda" else None


def sync(nat):
    if nat.device.type == "cuda":
        torch.cuda.synchronize()


def run_both(nat, dbl, name, tensors, out_names, *args, tol=2e-3, **kw):
    """tensors: dict name -> cpu tensor (inputs and pre-filled outputs).  Calls op `name` on both, compares outputs."""
    cpu = {k: (v.clone() if v is not None else None) for k, v in tensors.items()}
    # the B200 — or the host, when `nat` is the kernel source on the CUDA-on-CPU shim (tests/test_kernel_sources_cpu_shim.py reuses these
    # test bodies; there every buffer ends at an inaccessible page, so an out-of-bounds access is a crash)
    place = getattr(nat, "guarded", None) or (lambda v: v.clone().to(nat.device))
    gpu = {k: (place(v) if v is not None else None) for k, v in tensors.items()}

    def bind(o, t):
        a = [t[x] if isinstance(x, str) and x in t else x for x in args]
        k = {kk: (t[vv] if isinstance(vv, str) and vv in t else vv) for kk, vv in kw.items()}
        if "qkv" in k and k["qkv"] is not None:
            k["qkv"] = {qq: (t[qv] if isinstance(qv, str) else qv) for qq, qv in k["qkv"].items()}
        return getattr(o, name)(*a, **k)

    bind(dbl, cpu)(None)
    bind(nat, gpu)(stream_of(nat))
    sync(nat)
    for o in out_names:
        a, b = gpu[o].float().cpu(), cpu[o].float()
        assert torch.isfinite(a).all(), f"{name}: non-finite values in {o}"
        err = (a - b).abs().max().item()
        scale = b.abs().max().item() + 1e-6
        assert err <= tol * scale, f"{name}:{o} max err {err:.3e} vs scale {scale:.3e}"


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


@pytest.mark.parametrize("M,N,K", [(256, 320, 320), (1024, 1280, 5120), (130, 5, 2880), (2048, 768, 256), (32768, 320, 320),
                                   (16384, 640, 2560), (300, 130, 736)])
def test_gemm_plain_bias_residual(nat, dbl, M, N, K):
    t = {"A": rnd(M, K, dtype=torch.float16), "W": rnd(N, K, dtype=torch.float16, seed=1, scale=K ** -0.5),
         "bias": rnd(N, seed=2), "res": rnd(M, N, seed=3), "out": torch.zeros(M, N)}
    run_both(nat, dbl, "gemm", t, ["out"], "A", "W", "out", M, N, K, bias="bias", residual="res", ldr=N)


def test_gemm_splitk_rowbias(nat, dbl):
    M, N, K = 256, 1280, 11520
    t = {"A": rnd(M, K, dtype=torch.float16), "W": rnd(N, K, dtype=torch.float16, seed=1, scale=K ** -0.5),
         "rb": rnd(4, N, seed=2), "res": rnd(M, N, seed=3), "out": torch.zeros(M, N), "ws": torch.zeros(32 << 20, dtype=torch.uint8)}
    for split in (8, 0, 13):  # explicit, automatic, and a slice count that does not divide the k-blocks
        run_both(nat, dbl, "gemm", t, ["out"], "A", "W", "out", M, N, K, rowbias="rb", rows_per_group=64, residual="res", ldr=N,
                 split_k=split, ws="ws")


def test_gemm_geglu_f16(nat, dbl):
    M, C = 512, 320
    t = {"A": rnd(M, C, dtype=torch.float16), "W": rnd(8 * C, C, dtype=torch.float16, seed=1, scale=C ** -0.5),
         "bias": rnd(8 * C, seed=2, scale=0.1), "out": torch.zeros(M, 4 * C, dtype=torch.float16)}
    for tile in (256, 128):
        run_both(nat, dbl, "gemm", t, ["out"], "A", "W", "out", M, 8 * C, C, bias="bias", act=OPS.ACT_GEGLU, tile_n=tile, ldc=4 * C)


def test_gemm_gelu_colscale(nat, dbl):
    M, N, K = 4096, 256, 512
    t = {"A": rnd(M, K, dtype=torch.float16), "W": rnd(N, K, dtype=torch.float16, seed=1, scale=K ** -0.5),
         "bias": rnd(N, seed=2), "cs": rnd(N, seed=4), "res": rnd(M, N, seed=3), "out": torch.zeros(M, N)}
    run_both(nat, dbl, "gemm", t, ["out"], "A", "W", "out", M, N, K, bias="bias", colscale="cs", residual="res", ldr=N)
    t["out"] = torch.zeros(M, N, dtype=torch.float16)
    run_both(nat, dbl, "gemm", t, ["out"], "A", "W", "out", M, N, K, bias="bias", act=OPS.ACT_GELU)


@pytest.mark.parametrize("M,N,K,split", [(4096, 320, 320, 1), (1000, 640, 2560, 1), (256, 1280, 1280, 3), (130, 36, 96, 1)])
def test_gemm_fp16_copy_of_the_output(nat, dbl, M, N, K, split):
    """out16: the fp32 result once more as fp16, contiguous and as a column window of a wider (concatenation) buffer"""
    t = {"A": rnd(M, K, dtype=torch.float16), "W": rnd(N, K, dtype=torch.float16, seed=1, scale=K ** -0.5),
         "bias": rnd(N, seed=2), "res": rnd(M, N, seed=3), "out": torch.zeros(M, N), "o16": torch.zeros(M, N, dtype=torch.float16),
         "wide": torch.zeros(M, N + 64, dtype=torch.float16), "ws": torch.zeros(32 << 20, dtype=torch.uint8)}
    run_both(nat, dbl, "gemm", t, ["out", "o16"], "A", "W", "out", M, N, K, bias="bias", residual="res", ldr=N, out16="o16",
             split_k=split, ws="ws")
    cpu_w, gpu_w = t["wide"].clone(), t["wide"].cuda()
    args = dict(bias=t["bias"], residual=t["res"], ldr=N, ld16=N + 64)
    dbl.gemm(t["A"], t["W"], t["out"].clone(), M, N, K, out16=cpu_w[:, 64:], **args)(None)
    g = {k: v.cuda() for k, v in t.items()}
    nat.gemm(g["A"], g["W"], g["out"], M, N, K, bias=g["bias"], residual=g["res"], ldr=N, out16=gpu_w[:, 64:], ld16=N + 64)(
        torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (gpu_w[:, :64] == 0).all()
    assert (gpu_w.cpu().float() - cpu_w.float()).abs().max() <= 2e-3 * cpu_w.float().abs().max()


def _hilo16(x):
    hi = x.half()
    return hi, (x - hi.float()).half()


@pytest.mark.parametrize("M,N,K,split", [(2048, 320, 640, 1), (512, 1280, 2560, 3), (4096, 640, 960, 1), (300, 200, 128, 1)])
def test_gemm_split_precision_operands(nat, dbl, M, N, K, split):
    """hilo (ABI 9): A = [A_hi | A_lo], W = [W_hi | W_lo]; the result follows the fp32 operands to ~1e-6 instead of fp16's 5e-4,
    also when A is a column window of a wider [hi | lo] buffer (a_lo_off > K) and the rounding residual of the output is stored
    next to its fp16 copy (out16_lo)."""
    a32, w32 = rnd(M, K), rnd(N, K, seed=1, scale=K ** -0.5)
    ahi, alo = _hilo16(a32)
    whi, wlo = _hilo16(w32)
    pad = 192  # A as a window: columns [pad, pad + K) of a [M, 2 * (pad + K)] buffer, lo half pad + K further right
    wide = torch.zeros(M, 2 * (pad + K), dtype=torch.float16)
    wide[:, pad:pad + K], wide[:, 2 * pad + K:] = ahi, alo
    t = {"A": torch.cat([ahi, alo], 1), "W": torch.cat([whi, wlo], 1), "bias": rnd(N, seed=2), "out": torch.zeros(M, N),
         "o16": torch.zeros(M, 2 * N + 8, dtype=torch.float16), "ws": torch.zeros(32 << 20, dtype=torch.uint8)}
    # (o16 is compared as hi + lo below: a 1e-6 difference of the accumulators may move the fp16 rounding of a single value)
    run_both(nat, dbl, "gemm", t, ["out"], "A", "W", "out", M, N, K, bias="bias", hilo=True, out16="o16", ld16=2 * N + 8,
             out16_lo=N + 8, split_k=split, ws="ws", tol=2e-5)
    g = {k: v.cuda() for k, v in t.items()}
    out_w = torch.zeros(M, N, device="cuda")
    gw = wide.cuda()
    nat.gemm(gw[:, pad:], g["W"], out_w, M, N, K, bias=g["bias"], hilo=True, lda=2 * (pad + K), a_lo_off=pad + K)(
        torch.cuda.current_stream().cuda_stream)
    nat.gemm(g["A"], g["W"], g["out"], M, N, K, bias=g["bias"], hilo=True, out16=g["o16"], ld16=2 * N + 8, out16_lo=N + 8)(
        torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    exact = (a32.double() @ w32.double().t() + t["bias"].double()).float()
    assert (out_w.cpu() - exact).norm() / exact.norm() < 2e-5  # fp16 subnormal granularity of W_lo (2^-24 absolute) bounds it
    assert (g["out"].cpu() - exact).norm() / exact.norm() < 2e-5  # fp16 subnormal granularity of W_lo (2^-24 absolute) bounds it
    o16 = g["o16"].cpu().float()
    assert ((o16[:, :N] + o16[:, N + 8:]) - exact).norm() / exact.norm() < 2e-5  # fp16 subnormal granularity of W_lo (2^-24 absolute) bounds it  # hi + lo of the stored output


@pytest.mark.parametrize("n,H,C,Cout", [(2, 32, 320, 5), (3, 16, 128, 64)])
def test_conv3x3_split_precision(nat, dbl, n, H, C, Cout):
    """the UNet head (GroupNorm-SiLU-conv 320 -> 5) with [hi | lo] channels and [W_hi | W_lo] weights"""
    import torch.nn.functional as F
    M = n * H * H
    ldc = 8 if Cout == 5 else Cout
    a32, w32 = rnd(M, C), rnd(Cout, 9 * C, seed=1, scale=(9 * C) ** -0.5)
    t = {"A": torch.cat(_hilo16(a32), 1), "W": torch.cat(_hilo16(w32), 1), "bias": rnd(Cout, seed=2), "out": torch.zeros(M, ldc),
         "ws": torch.zeros(32 << 20, dtype=torch.uint8)}
    run_both(nat, dbl, "gemm", t, ["out"], "A", "W", "out", M, Cout, 9 * C, conv=(n, H, H, C), bias="bias", ldc=ldc, hilo=True,
             split_k=1, tol=2e-5)
    g = {k: v.cuda() for k, v in t.items()}
    nat.gemm(g["A"], g["W"], g["out"], M, Cout, 9 * C, conv=(n, H, H, C), bias=g["bias"], ldc=ldc, hilo=True)(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    x = a32.reshape(n, H, H, C).permute(0, 3, 1, 2).double()
    w = w32.reshape(Cout, 3, 3, C).permute(0, 3, 1, 2).double()
    exact = F.conv2d(x, w, t["bias"].double(), padding=1).permute(0, 2, 3, 1).reshape(M, Cout).float()
    assert (g["out"].cpu()[:, :Cout] - exact).norm() / exact.norm() < 2e-5  # fp16 subnormal granularity of W_lo (2^-24 absolute) bounds it


@pytest.mark.parametrize("n,hw,C", [(2, 1024, 320), (3, 4096, 64), (16, 1024, 320)])
def test_groupnorm_split_precision_output(nat, dbl, n, hw, C):
    t = {"x": rnd(n * hw, C) * 2 + 0.5, "g": rnd(C, seed=1), "b": rnd(C, seed=2), "y": torch.zeros(n * hw, 2 * C, dtype=torch.float16)}
    run_both(nat, dbl, "groupnorm_hilo", t, ["y"], "x", "g", "b", "y", n, hw, C, 1e-5, True)
    y = nat.empty((n * hw, 2 * C), torch.float16)
    dev = nat.device
    place = getattr(nat, "guarded", None) or (lambda v: v.clone().to(dev))
    nat.groupnorm_hilo(place(t["x"]), place(t["g"]), place(t["b"]), y, n, hw, C, 1e-5, True)(stream_of(nat))
    sync(nat)
    import torch.nn.functional as F
    exact = F.silu(F.group_norm(t["x"].double().reshape(n, hw, C).permute(0, 2, 1), 32, t["g"].double(), t["b"].double(), 1e-5))
    exact = exact.permute(0, 2, 1).reshape(n * hw, C).float()
    got = y.cpu().float()
    assert ((got[:, :C] + got[:, C:]) - exact).norm() / exact.norm() < 5e-6


@pytest.mark.parametrize("n,H,Cin,Cout", [(2, 32, 320, 320), (4, 16, 640, 320), (16, 4, 1280, 1280), (3, 8, 960, 640), (2, 32, 16, 320),
                                          (2, 32, 320, 5)])
def test_conv3x3(nat, dbl, n, H, Cin, Cout):
    M = n * H * H
    ldc = 8 if Cout == 5 else Cout
    t = {"A": rnd(M, Cin, dtype=torch.float16), "W": rnd(Cout, 9 * Cin, dtype=torch.float16, seed=1, scale=(9 * Cin) ** -0.5),
         "bias": rnd(Cout, seed=2), "out": torch.zeros(M, ldc), "ws": torch.zeros(32 << 20, dtype=torch.uint8)}
    run_both(nat, dbl, "gemm", t, ["out"], "A", "W", "out", M, Cout, 9 * Cin, conv=(n, H, H, Cin), bias="bias", ldc=ldc,
             split_k=(0 if H <= 8 else 1), ws="ws")


@pytest.mark.parametrize("n,H,C,Cout,tile,split,pair", [(16, 4, 1280, 1280, 0, 0, 0), (16, 4, 1280, 1280, 160, 2, 2), (16, 16, 640, 640, 160, 1, 2),
                                                        (2, 16, 640, 640, 0, 1, 0), (3, 8, 64, 32, 0, 1, 1), (16, 4, 1280, 1280, 256, 3, 1),
                                                        (2, 4, 1280, 1280, 256, 7, 1)])
def test_conv3x3_behind_nearest_upsample_as_phase_convolutions(nat, dbl, n, H, C, Cout, tile, split, pair):
    """ABI 14 (conv_up2): Upsample = nearest x2 + conv3x3 as four 2 x 2 phase convolutions of the source image in one launch — against
    the emulation and against F.interpolate + F.conv2d in fp32; output rows interleave the phases, out16 into a column window"""
    import torch.nn.functional as F
    from mvdfusion_b200.engine import PackedWeights
    g = torch.Generator().manual_seed(n * 1000 + H)
    w = torch.randn(Cout, C, 3, 3, generator=g) * (9 * C) ** -0.5
    bias = torch.randn(Cout, generator=g)
    x = torch.randn(n, H, H, C, generator=g)
    W = PackedWeights({"k.weight": w, "k.bias": bias}, dbl)
    wp, bp = W.conv3_up2("k.weight"), W.f32_tiled("k.bias", 4)
    M, Mo = n * H * H, n * 4 * H * H
    x16 = x.half().reshape(M, C)
    outs = {}
    for side, ops, dev in (("cpu", dbl, "cpu"), ("gpu", nat, "cuda")):
        out = torch.zeros(Mo, Cout, device=dev)
        wide = torch.zeros(Mo, Cout + 64, dtype=torch.float16, device=dev)
        kw = dict(tile_n=tile, split_k=split, cta_pair=pair, ws=torch.zeros(64 << 20, dtype=torch.uint8, device=dev)) if side == "gpu" else {}
        ops.gemm(x16.to(dev), wp.to(dev), out, M, 4 * Cout, 4 * C, conv=(n, H, H, C), conv_up2=True, bias=bp.to(dev), ldc=Cout,
                 out16=wide[:, 64:], ld16=Cout + 64, **kw)(torch.cuda.current_stream().cuda_stream if side == "gpu" else None)
        outs[side] = (out, wide)
    torch.cuda.synchronize()
    _close(outs["gpu"][0], outs["cpu"][0], 2e-3, "phase convolutions")
    _close(outs["gpu"][1], outs["cpu"][1], 2e-3, "fp16 copy")
    assert (outs["gpu"][1][:, :64] == 0).all()
    ref = F.conv2d(F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="nearest"), w, bias, padding=1).permute(0, 2, 3, 1).reshape(Mo, Cout)
    assert rel(outs["gpu"][0].cpu(), ref) < 1e-3


def _close(a, b, tol, what):
    a, b = a.float().cpu(), b.float().cpu()
    assert torch.isfinite(a).all(), f"{what}: non-finite values"
    err, scale = (a - b).abs().max().item(), b.abs().max().item() + 1e-6
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("M,C,res,tile,pair", [(4096, 320, True, 0, 0), (4096, 320, False, 320, 2), (2048, 320, True, 160, 1), (1000, 640, True, 0, 0),
                                               (512, 1280, True, 96, 1), (256, 1280, False, 32, 1), (130, 64, True, 0, 0), (16384, 320, True, 320, 2)])
def test_gemm_layernorm_statistics_from_the_epilogue(nat, dbl, M, C, res, tile, pair):
    """ABI 13, producer side: next to the fp32 rows and their fp16 copy the epilogue leaves (sum, sum of squares) per row and
    32-column chunk — every tile geometry the step uses for the C x C products (single CTAs, pairs, 320-column pair tiles, M tails)"""
    t = {"A": rnd(M, C, dtype=torch.float16), "W": rnd(C, C, dtype=torch.float16, seed=1, scale=C ** -0.5), "bias": rnd(C, seed=2),
         "rb": rnd((M + 255) // 256, C, seed=4), "res": rnd(M, C, seed=3, scale=3.0) + 0.7}
    outs = {}
    for side, ops, dev in (("cpu", dbl, "cpu"), ("gpu", nat, "cuda")):
        g = {k: v.to(dev) for k, v in t.items()}
        o = {"out": torch.zeros(M, C, device=dev), "o16": torch.zeros(M, C, dtype=torch.float16, device=dev),
             "st": torch.full((C // 32, M, 2), float("nan"), device=dev)}
        kw = dict(bias=g["bias"], rowbias=g["rb"], rows_per_group=256, out16=o["o16"], ld16=C, ln_stats_out=o["st"])
        if res:
            kw.update(residual=g["res"], ldr=C)
        if side == "gpu":
            kw.update(tile_n=tile, cta_pair=pair)
        ops.gemm(g["A"], g["W"], o["out"], M, C, C, **kw)(torch.cuda.current_stream().cuda_stream if side == "gpu" else None)
        outs[side] = o
    torch.cuda.synchronize()
    _close(outs["gpu"]["out"], outs["cpu"]["out"], 2e-3, "out")
    _close(outs["gpu"]["o16"], outs["cpu"]["o16"], 2e-3, "out16")
    _close(outs["gpu"]["st"][..., 0], outs["cpu"]["st"][..., 0], 2e-3, "chunk sums")
    _close(outs["gpu"]["st"][..., 1], outs["cpu"]["st"][..., 1], 2e-3, "chunk sums of squares")
    # the statistics describe the rows that were written (not the accumulators): mean / variance from them == those of `out`
    x = outs["gpu"]["out"].double().cpu()
    st = outs["gpu"]["st"].double().cpu().sum(0)
    mu = st[:, 0] / C
    var = st[:, 1] / C - mu * mu
    assert (mu - x.mean(1)).abs().max() < 1e-5 * (1 + x.abs().max())
    assert ((var - x.var(1, unbiased=False)).abs() / x.var(1, unbiased=False)).max() < 1e-4


@pytest.mark.parametrize("n,seq,C,pair", [(2, 1024, 320, 0), (2, 256, 640, 2), (3, 64, 1280, 0), (2, 16, 1280, 1), (1, 200, 64, 0), (16, 1024, 320, 0)])
def test_gemm_layernorm_folded_into_qkv_and_geglu(nat, dbl, n, seq, C, pair):
    """ABI 13, consumer side: raw fp16 rows + their chunk statistics in, LayerNorm(x) W^T (+ b) out — the QKV head scatter and the
    GEGLU epilogue, against the emulation and against nn.LayerNorm followed by the plain products in fp32"""
    heads = 8
    d = C // heads
    dpad = (d + 63) // 64 * 64
    M = n * seq
    nb = n * heads * seq * dpad
    x = rnd(M, C, seed=11, scale=2.5) + 0.9 * rnd(M, 1, seed=12)
    gamma, beta = 1.0 + 0.3 * rnd(C, seed=13), 0.2 * rnd(C, seed=14)
    ch = x.reshape(M, C // 32, 32)
    st = torch.stack([ch.sum(-1), (ch * ch).sum(-1)], dim=-1).permute(1, 0, 2).contiguous()
    x16 = x.half()
    ln = torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-5)
    stream = lambda side: torch.cuda.current_stream().cuda_stream if side == "gpu" else None
    # --- QKV
    Wq = rnd(3 * C, C, seed=1, scale=C ** -0.5)
    wq = (Wq * gamma[None, :]).half()
    cs, bq = wq.float().sum(1), Wq @ beta
    got = {}
    for side, ops, dev in (("cpu", dbl, "cpu"), ("gpu", nat, "cuda")):
        q, k, vt = (torch.zeros(nb, dtype=torch.float16, device=dev) for _ in range(3))
        kw = dict(cta_pair=pair) if side == "gpu" else {}
        ops.gemm(x16.to(dev), wq.to(dev), q, M, 3 * C, C, bias=bq.to(dev), ln=(st.to(dev), cs.to(dev), 1e-5),
                 qkv=dict(out_k=k, out_vt=vt, heads=heads, dhead=d, dpad=dpad, seq=seq), **kw)(stream(side))
        got[side] = (q, k, vt)
    torch.cuda.synchronize()
    for a, b, nm in zip(got["gpu"], got["cpu"], "q k vt".split()):
        _close(a, b, 2e-3, "folded QKV " + nm)
    ref = (ln @ Wq.t()).reshape(n, seq, 3, heads, d).permute(2, 0, 3, 1, 4)
    qg = got["gpu"][0].reshape(n, heads, seq, dpad)[..., :d].float().cpu()
    vg = got["gpu"][2].reshape(n, heads, dpad, seq)[:, :, :d].float().cpu().transpose(-1, -2)
    assert rel(qg, ref[0]) < 2e-3 and rel(vg, ref[2]) < 2e-3
    # --- GEGLU
    from ops_double import geglu_permutation
    tile = 256
    Wg, bg = rnd(8 * C, C, seed=2, scale=C ** -0.5), rnd(8 * C, seed=3, scale=0.1)
    perm = geglu_permutation(4 * C, tile)
    wg = (Wg[perm] * gamma[None, :]).half()
    csg, bgp = wg.float().sum(1), (Wg @ beta + bg)[perm]
    outs = {}
    for side, ops, dev in (("cpu", dbl, "cpu"), ("gpu", nat, "cuda")):
        o = torch.zeros(M, 4 * C, dtype=torch.float16, device=dev)
        kw = dict(cta_pair=pair) if side == "gpu" else {}
        ops.gemm(x16.to(dev), wg.to(dev), o, M, 8 * C, C, bias=bgp.to(dev), act=OPS.ACT_GEGLU, tile_n=tile, ldc=4 * C,
                 ln=(st.to(dev), csg.to(dev), 1e-5), **kw)(stream(side))
        outs[side] = o
    torch.cuda.synchronize()
    _close(outs["gpu"], outs["cpu"], 3e-3, "folded GEGLU")
    y = ln @ Wg.t() + bg
    assert rel(outs["gpu"].float().cpu(), y[:, :4 * C] * torch.nn.functional.gelu(y[:, 4 * C:])) < 3e-3


def test_gemm_layernorm_arguments_are_checked(nat):
    M, C = 256, 320
    dev = "cuda"
    A, W = torch.zeros(M, C, dtype=torch.float16, device=dev), torch.zeros(C, C, dtype=torch.float16, device=dev)
    st = torch.zeros(C // 32, M, 2, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    with pytest.raises(OPS.MvdError):  # statistics need the TMA epilogue: no split-K
        nat.gemm(A, W, torch.zeros(M, C, device=dev), M, C, C, ln_stats_out=st, split_k=2, ws=torch.zeros(1 << 24, dtype=torch.uint8, device=dev))(s)
    with pytest.raises(OPS.MvdError):  # ... and an fp32 output
        nat.gemm(A, W, torch.zeros(M, C, dtype=torch.float16, device=dev), M, C, C, ln_stats_out=st)(s)
    with pytest.raises(OPS.MvdError):  # the fold exists for the QKV scatter and GEGLU only
        nat.gemm(A, W, torch.zeros(M, C, device=dev), M, C, C, ln=(st, torch.zeros(C, device=dev), 1e-5))(s)


@pytest.mark.parametrize("n,seq,C", [(2, 1024, 320), (2, 256, 640), (3, 64, 1280), (2, 16, 1280), (2, 1024, 64)])
def test_qkv_and_attention(nat, dbl, n, seq, C):
    heads = 8
    d = C // heads
    dpad = (d + 63) // 64 * 64
    M = n * seq
    nb = n * heads * seq * dpad
    t = {"A": rnd(M, C, dtype=torch.float16), "W": rnd(3 * C, C, dtype=torch.float16, seed=1, scale=C ** -0.5),
         "q": torch.zeros(nb, dtype=torch.float16), "k": torch.zeros(nb, dtype=torch.float16),
         "vt": torch.zeros(nb, dtype=torch.float16), "o": torch.zeros(M, C, dtype=torch.float16)}
    run_both(nat, dbl, "gemm", t, ["q", "k", "vt"], "A", "W", "q", M, 3 * C, C,
             qkv=dict(out_k="k", out_vt="vt", heads=heads, dhead=d, dpad=dpad, seq=seq))
    t["bias"] = rnd(3 * C, seed=5)  # bias on the scattered output: staged q / k chunks and the direct v^T chunks
    run_both(nat, dbl, "gemm", t, ["q", "k", "vt"], "A", "W", "q", M, 3 * C, C, bias="bias",
             qkv=dict(out_k="k", out_vt="vt", heads=heads, dhead=d, dpad=dpad, seq=seq))
    del t["bias"]
    # attention on the (emulated) q/k/v so that both sides see identical operands
    from ops_double import TorchOpsDouble
    d2 = TorchOpsDouble()
    d2.gemm(t["A"], t["W"], t["q"], M, 3 * C, C, qkv=dict(out_k=t["k"], out_vt=t["vt"], heads=heads, dhead=d, dpad=dpad, seq=seq))(None)
    run_both(nat, dbl, "attn_self", t, ["o"], "q", "k", "vt", "o", n, heads, seq, d, dpad, C, tol=4e-3)


@pytest.mark.parametrize("n,hw,C,silu", [(2, 1024, 320, True), (16, 16, 2560, True), (3, 256, 1920, False), (2, 1024, 960, True),
                                         (2, 4096, 320, True), (1, 4096, 960, False), (3, 64, 1280, True), (2, 256, 640, True),
                                         (1, 100, 64, True), (2, 16384, 32, False),
                                         # enough images for the slab form (a cluster of 8 slabs per image held in shared memory)
                                         (16, 1024, 320, True), (16, 256, 640, True), (16, 64, 1280, True), (16, 16, 1280, True), (12, 256, 1280, False)])
def test_groupnorm(nat, dbl, n, hw, C, silu):
    t = {"x": rnd(n * hw, C) * 2 + 0.5, "g": rnd(C, seed=1), "b": rnd(C, seed=2), "y": torch.zeros(n * hw, C, dtype=torch.float16),
         "ws": torch.zeros(max(n, 64) * 64, dtype=torch.float64)}
    run_both(nat, dbl, "groupnorm", t, ["y"], "x", "g", "b", "y", "ws", n, hw, C, 1e-5, silu)


@pytest.mark.parametrize("n,hw,C1,C2", [(2, 1024, 320, 320), (3, 256, 640, 320), (4, 16, 1280, 1280), (2, 64, 1280, 640),
                                        (16, 256, 320, 320), (16, 64, 640, 640), (16, 16, 640, 640), (16, 256, 640, 320)])
def test_groupnorm_two_sources_and_concat16(nat, dbl, n, hw, C1, C2):
    C = C1 + C2
    t = {"a": rnd(n * hw, C1, seed=1) + 0.3, "b": rnd(n * hw, C2, seed=2) * 2.0, "g": rnd(C, seed=3), "be": rnd(C, seed=4),
         "y": torch.zeros(n * hw, C, dtype=torch.float16), "c": torch.zeros(n * hw, C, dtype=torch.float16)}
    run_both(nat, dbl, "groupnorm2", t, ["y"], "a", C1, "b", C2, "g", "be", "y", n, hw, 1e-5, True, tol=4e-3)
    run_both(nat, dbl, "concat16", t, ["c"], "a", "b", "c", n * hw, C1, C2, tol=1e-3)


def test_layernorms(nat, dbl):
    for rows, C in ((512, 320), (100, 1280), (4096, 256)):
        t = {"x": rnd(rows, C) * 3, "g": rnd(C, seed=1), "b": rnd(C, seed=2), "y": torch.zeros(rows, C, dtype=torch.float16)}
        run_both(nat, dbl, "layernorm", t, ["y"], "x", "g", "b", "y", rows, C, 1e-5)
        run_both(nat, dbl, "ln_modulate", t, ["y"], "x", "b", "g", "y", rows, C, 1e-6)


def test_data_movement(nat, dbl):
    n, H, C = 3, 8, 64
    t = {"x": rnd(n * H * H, C), "y": torch.zeros(n * 4 * H * H, C, dtype=torch.float16)}
    run_both(nat, dbl, "upsample2x", t, ["y"], "x", "y", n, H, H, C)
    t = {"x": rnd(n * H * H, C), "y": torch.zeros(n * H * H // 4, 9 * C, dtype=torch.float16)}
    run_both(nat, dbl, "im2col_s2", t, ["y"], "x", "y", n, H, H, C)
    t = {"a": rnd(100, 64), "b": rnd(100, 32, seed=1), "o": torch.zeros(100, 96)}
    run_both(nat, dbl, "concat", t, ["o"], "a", "b", "o", 100, 64, 32)
    t = {"x": rnd(1000), "y": torch.zeros(1000, dtype=torch.float16)}
    run_both(nat, dbl, "cast", t, ["y"], "x", "y", 1000)
    t = {"x": rnd(2, 10, 64), "y": torch.zeros(2 * 64, 16, dtype=torch.float16), "r": torch.zeros(2 * 64, 10), "z": torch.zeros(2, 10, 64)}
    run_both(nat, dbl, "nchw_to_nhwc16", t, ["y"], "x", "y", 2, 10, 64, 16)
    run_both(nat, dbl, "nchw_to_rows", t, ["r"], "x", "r", 2, 10, 64)
    t["r"] = rnd(2 * 64, 10)
    run_both(nat, dbl, "rows_to_nchw", t, ["z"], "r", "z", 2, 10, 10, 64)


@pytest.mark.parametrize("P,V", [(1030, 8), (515, 4), (129, 16), (200, 3)])
def test_view_attention_layouts(nat, dbl, P, V):
    """staged kernel (heads * V divides 256, ragged last CTA) and the per-thread fallback (V = 3)"""
    t = {"qkv": rnd(P * V, 768, dtype=torch.float16, seed=3), "o": torch.zeros(P * V, 256, dtype=torch.float16)}
    run_both(nat, dbl, "view_attention", t, ["o"], "qkv", "o", P, V, 8, 32)


def test_gemv_grouped(nat, dbl):
    K = 1280
    x = rnd(1, K, seed=5)
    sizes = (320, 640, 1280, 40, 1280)
    Ws = [rnd(n, K, dtype=torch.float16, seed=10 + i, scale=K ** -0.5) for i, n in enumerate(sizes)]
    bs = [rnd(n, seed=20 + i) if i != 3 else None for i, n in enumerate(sizes)]
    y_cpu = [torch.zeros(n) for n in sizes]
    dev = nat.device
    place = getattr(nat, "guarded", None) or (lambda v: v.clone().to(dev))
    y_gpu = [place(torch.zeros(n)) for n in sizes]
    dbl.gemv_grouped(x, K, list(zip(Ws, bs, y_cpu)), silu_in=True)(None)
    nat.gemv_grouped(place(x), K, [(place(w), place(b) if b is not None else None, y) for w, b, y in zip(Ws, bs, y_gpu)],
                     silu_in=True)(stream_of(nat))
    sync(nat)
    for a, b in zip(y_gpu, y_cpu):
        assert (a.cpu() - b).abs().max().item() <= 1e-4 * (b.abs().max().item() + 1e-6)


def test_gemv_and_timestep(nat, dbl):
    for M, N, K in ((1, 1280, 320), (8, 768, 796), (1, 1536, 256)):
        k8 = (K + 7) // 8 * 8
        W = torch.zeros(N, k8, dtype=torch.float16)
        W[:, :K] = rnd(N, K, dtype=torch.float16, scale=K ** -0.5)
        t = {"x": rnd(M, K, seed=1), "W": W, "b": rnd(N, seed=2), "y": torch.zeros(M, N)}
        run_both(nat, dbl, "gemv", t, ["y"], "x", "W", "b", "y", M, N, K, ldx=K, ldw=k8, ldy=N, silu_in=True, silu_out=True, tol=1e-4)
    import math
    freqs = torch.exp(-math.log(10000) * torch.arange(160, dtype=torch.float32) / 160)
    t = {"t": torch.tensor([981.0]), "f": freqs, "o": torch.zeros(320)}
    run_both(nat, dbl, "timestep_embedding", t, ["o"], "t", "f", "o", 320, tol=2e-4)


def test_unet_input_cfg_ddim_tables(nat, dbl):
    n, hw = 3, 64
    t = {"noisy": rnd(n, 5, hw), "cond": rnd(1, 5, hw, seed=1), "cs": torch.tensor([1.0, 0.0, 1.0]),
         "out": torch.zeros(2 * n * hw, 16, dtype=torch.float16)}
    run_both(nat, dbl, "unet_input", t, ["out"], "noisy", "cond", False, "cs", "out", n, 2 * n, hw, 16)
    t32 = dict(t, out=torch.zeros(2 * n * hw, 32, dtype=torch.float16))  # [hi | lo | hi | 0] channels of the split-precision stem
    run_both(nat, dbl, "unet_input", t32, ["out"], "noisy", "cond", False, "cs", "out", n, 2 * n, hw, 32, hilo=True, tol=0)
    coef = torch.tensor([0.5, 0.6, 0.70710678, 0.2, 1.0, 2.5])
    t = {"head": rnd(2 * n * hw, 8), "coef": coef, "xt": rnd(n, 5, hw, seed=2), "noise": rnd(n, 5, hw, seed=3),
         "eps": torch.zeros(n, 5, hw), "xp": torch.zeros(n, 5, hw), "x0": torch.zeros(n, 5, hw)}
    run_both(nat, dbl, "cfg_ddim", t, ["eps", "xp", "x0"], "head", 8, True, "coef", "xt", "noise", "eps", "xp", "x0", n, hw, tol=1e-5)
    t = {"tab": rnd(7, 33), "idx": torch.tensor([4], dtype=torch.int32), "o": torch.zeros(33)}
    run_both(nat, dbl, "gather_rows", t, ["o"], "tab", 33, "idx", "o", tol=0)
    run_both(nat, dbl, "increment", t, ["idx"], "idx", 2, tol=0)


@pytest.mark.parametrize("N,D,q_first,q_count", [(4, 1, 0, 4), (3, 3, 1, 2)])
def test_gridattn_kernels(nat, dbl, N, D, q_first, q_count):
    from mvdfusion_b200 import synthetic
    from mvdfusion_b200.denoise import pack_cameras
    S, hw = 32, 1024
    R, T, f, p = synthetic.gso_rig(N)
    cams = pack_cameras(torch.cat([R[1:], R[:1]]), torch.cat([T[1:], T[:1]]), torch.cat([f[1:], f[:1]]), torch.cat([p[1:], p[:1]]))
    half = 1.0 / S
    t = {"noisy": rnd(N, 5, hw), "inp": rnd(1, 5, hw, seed=1), "eps": rnd(N, D, hw, seed=2), "scal": torch.tensor([0.6, 0.13]),
         "Wz": rnd(256 * 5, seed=3, scale=0.4), "bz": rnd(256, seed=4, scale=0.1),
         "feat": torch.zeros((N + 1) * hw, 256, dtype=torch.float16), "z": torch.zeros(N * D * hw)}
    run_both(nat, dbl, "gridattn_prep", t, ["feat", "z"], "noisy", "inp", None, "eps", "scal", "Wz", "bz", "feat", "z", N, S, D, 2.0, 0.5)
    from ops_double import TorchOpsDouble
    TorchOpsDouble().gridattn_prep(t["noisy"], t["inp"], None, t["eps"], t["scal"], t["Wz"], t["bz"], t["feat"], t["z"], N, S, D, 2.0, 0.5)(None)
    P = q_count * hw * D
    t.update({"cams": cams, "mask": torch.ones(N), "fr": (2.0 ** torch.arange(7, dtype=torch.float32)) * 0.1,
              "grid": torch.linspace(1.0 - half, -1.0 + half, S), "tok": torch.zeros(P * N, 736, dtype=torch.float16)})
    run_both(nat, dbl, "gridattn_tokens", t, ["tok"], "feat", "z", "cams", "mask", "fr", "grid", "tok", N, S, D, q_first, q_count, tol=4e-3)
    t = {"qkv": rnd(P * N, 768, dtype=torch.float16), "o": torch.zeros(P * N, 256, dtype=torch.float16)}
    run_both(nat, dbl, "view_attention", t, ["o"], "qkv", "o", P, N, 8, 32)
    t = {"x": rnd(P * N, 256), "w": rnd(256, seed=1, scale=0.1), "b": torch.tensor([0.1]), "o": torch.zeros(P, 256, dtype=torch.float16)}
    run_both(nat, dbl, "view_pool", t, ["o"], "x", "w", "b", "o", P, N, 256)
    t = {"i": rnd(2 * hw * D, 768, dtype=torch.float16), "o": torch.zeros(2 * (hw // 16) * D, 768, dtype=torch.float16)}
    run_both(nat, dbl, "frustum_pool", t, ["o"], "i", "o", 2, S, D, 768, 4)
    if D > 1:
        M, C = 512, 640
        t = {"q": rnd(M, C, dtype=torch.float16), "kv": rnd(M * D, 2 * C, dtype=torch.float16, seed=1), "o": torch.zeros(M, C, dtype=torch.float16)}
        run_both(nat, dbl, "pixel_cross_attn", t, ["o"], "q", "kv", "o", M, D, 8, C // 8)


@pytest.mark.parametrize("n,H,C,Cout,no_pad_lo,window", [(2, 16, 320, 320, False, False), (16, 4, 1280, 1280, False, True), (3, 8, 128, 64, True, False),
                                                          (2, 32, 64, 128, False, True)])
def test_conv3x3_stride2_implicit(nat, dbl, n, H, C, Cout, no_pad_lo, window):
    """ABI 10: Downsample.op as a strided implicit GEMM (TMA element strides 2) on the full-resolution fp16 image — dense, or a column
    window of a wider buffer (pixel pitch); H is the OUTPUT side"""
    M = n * H * H
    pitch = C + 192 if window else C
    img = rnd(n * 4 * H * H, pitch, dtype=torch.float16)
    t = {"A": img, "W": rnd(Cout, 9 * C, dtype=torch.float16, seed=1, scale=(9 * C) ** -0.5), "bias": rnd(Cout, seed=2), "out": torch.zeros(M, Cout),
         "ws": torch.zeros(32 << 20, dtype=torch.uint8)}
    cpu, gpu = {k: v.clone() for k, v in t.items()}, {k: v.cuda() for k, v in t.items()}
    off = 64 if window else 0
    kw = dict(conv=(n, H, H, C), conv_stride=2, conv_no_pad_lo=no_pad_lo, lda=pitch if window else None)
    dbl.gemm(cpu["A"][:, off:], cpu["W"], cpu["out"], M, Cout, 9 * C, bias=cpu["bias"], **kw)(None)
    nat.gemm(gpu["A"][:, off:], gpu["W"], gpu["out"], M, Cout, 9 * C, bias=gpu["bias"], split_k=(0 if H <= 8 else 1), ws=gpu["ws"], **kw)(
        torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err = (gpu["out"].cpu() - cpu["out"]).abs().max().item()
    assert err <= 2e-3 * cpu["out"].abs().max().item(), err
    # and against torch's strided convolution on the same fp16-rounded operands
    import torch.nn.functional as F
    x = img[:, off:off + C].float().reshape(n, 2 * H, 2 * H, C).permute(0, 3, 1, 2)
    w = t["W"].float().reshape(Cout, 3, 3, C).permute(0, 3, 1, 2)
    ref = F.conv2d(F.pad(x, (0, 1, 0, 1)) if no_pad_lo else x, w, t["bias"], stride=2, padding=0 if no_pad_lo else 1)
    assert rel(gpu["out"].cpu(), ref.permute(0, 2, 3, 1).reshape(M, Cout)) < 1e-3


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def test_attention_with_masked_key_padding(nat, dbl):
    """mvd_attn_self_masked_f16: CLIP's 257 tokens in a 272-row layout — keys [257, 272) never contribute"""
    n, heads, seq, valid, d = 2, 16, 272, 257, 64
    g = torch.Generator().manual_seed(3)
    q, k = torch.randn(n * heads, seq, d, generator=g).half(), torch.randn(n * heads, seq, d, generator=g).half()
    v = torch.randn(n * heads, seq, d, generator=g).half()
    k[:, valid:] = 50.0   # poison: huge scores if the padding keys were read
    v[:, valid:] = 100.0
    vt = v.transpose(1, 2).contiguous()
    t = {"q": q, "k": k, "vt": vt, "out": torch.zeros(n * seq, heads * d, dtype=torch.float16)}
    run_both(nat, dbl, "attn_self", t, ["out"], "q", "k", "vt", "out", n, heads, seq, d, 64, heads * d, seq_valid=valid, tol=4e-3)
    want = torch.softmax((q[:, :, None, :].float() * k[:, None, :valid].float()).sum(-1) * d ** -0.5, -1) @ v[:, :valid].float()
    out = nat.zeros((n * seq, heads * d), torch.float16)
    nat.attn_self(q.cuda(), k.cuda(), vt.cuda(), out, n, heads, seq, d, 64, heads * d, seq_valid=valid)(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().float().reshape(n, seq, heads, d).permute(0, 2, 1, 3).reshape(n * heads, seq, d)
    assert rel(got[:, :valid], want[:, :valid]) < 2e-3


def test_layernorm_f32_with_row_pitch(nat, dbl):
    rows, C, pitch = 5, 1024, 272 * 1024
    t = {"x": rnd(rows * pitch // 1024, 1024) * 3 + 1, "g": rnd(C, seed=1), "b": rnd(C, seed=2), "y": torch.zeros(rows, C)}
    run_both(nat, dbl, "layernorm_f32", t, ["y"], "x", "g", "b", "y", rows, C, 1e-5, ldx=pitch, tol=1e-5)
    t2 = {"x": rnd(300, 64), "g": rnd(64, seed=1), "b": rnd(64, seed=2), "y": torch.zeros(300, 64)}
    run_both(nat, dbl, "layernorm_f32", t2, ["y"], "x", "g", "b", "y", 300, 64, 1e-5, tol=1e-5)


def _dit_inputs(R, V, nl, token_k=736, seed=0):
    g = torch.Generator().manual_seed(100 + seed)
    r = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale
    t = {"tokens": r(R, token_k).half(), "w_pre": r(256, token_k, scale=token_k ** -0.5).half(), "b_pre": r(256, scale=0.1),
         "pool_w": r(256, scale=0.1), "pool_b": r(1), "pooled": torch.zeros(R // V, 256, dtype=torch.float16), "x_out": torch.zeros(R, 256)}
    layers = []
    for i in range(nl):
        lay = {"w_qkv": r(768, 256, scale=1 / 16).half(), "b_qkv": r(768, scale=0.1), "w_proj": r(256, 256, scale=1 / 32).half(), "b_proj": r(256, scale=0.1),
               "w_fc1": r(512, 256, scale=1 / 16).half(), "b_fc1": r(512, scale=0.1), "w_fc2": r(256, 512, scale=1 / 45).half(), "b_fc2": r(256, scale=0.1),
               "shift_msa": r(256, scale=0.2), "scale_msa": r(256, scale=0.2), "shift_mlp": r(256, scale=0.2), "scale_mlp": r(256, scale=0.2)}
        layers.append(lay)
    return t, layers


@pytest.mark.parametrize("R,V,nl", [(384, 8, 1), (384, 8, 3), (160, 8, 3), (1024, 16, 3), (512, 4, 2), (128 * 150 + 64, 8, 3), (256, 1, 1)])
def test_gridattn_dit_kernel(nat, dbl, R, V, nl):
    """mvd_gridattn_dit_f16 (pre_layer_b + DiT blocks + view pooling in one kernel) against the per-op emulation: ragged last tile,
    several tiles per CTA (150 tiles + a tail on 148 SMs), V = 1 / 4 / 8 / 16"""
    t, layers = _dit_inputs(R, V, nl)
    cpu_l, gpu_l = [dict(l) for l in layers], [{k: v.cuda() for k, v in l.items()} for l in layers]
    cpu, gpu = {k: v.clone() for k, v in t.items()}, {k: v.cuda() for k, v in t.items()}
    args = lambda d, ls: (d["tokens"], 736, d["w_pre"], d["b_pre"], ls, d["pool_w"], d["pool_b"], d["pooled"], R, V, 1e-6)
    dbl.gridattn_dit(*args(cpu, cpu_l), x_out=cpu["x_out"])(None)
    nat.gridattn_dit(*args(gpu, gpu_l), x_out=gpu["x_out"])(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for name, tol in (("x_out", 3e-3), ("pooled", 3e-3)):
        a, b = gpu[name].float().cpu(), cpu[name].float()
        assert torch.isfinite(a).all(), name
        assert float((a - b).norm() / b.norm()) < tol, (name, float((a - b).norm() / b.norm()), float((a - b).abs().max()))


def test_dit_fold_gates(nat, dbl):
    W1, W2 = rnd(256, 256, dtype=torch.float16), rnd(256, 512, dtype=torch.float16, seed=1)
    t = {"w1": W1, "w2": W2, "g1": rnd(256, seed=2), "g2": rnd(256, seed=3), "b1": rnd(256, seed=4), "b2": rnd(256, seed=5),
         "o1": torch.zeros_like(W1), "o2": torch.zeros_like(W2), "c1": torch.zeros(256), "c2": torch.zeros(256)}
    cpu, gpu = {k: v.clone() for k, v in t.items()}, {k: v.cuda() for k, v in t.items()}
    jobs = lambda d: [(d["w1"], d["g1"], d["b1"], d["o1"], d["c1"]), (d["w2"], d["g2"], d["b2"], d["o2"], d["c2"])]
    dbl.dit_fold_gates(jobs(cpu))(None)
    nat.dit_fold_gates(jobs(gpu))(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for k in ("o1", "o2", "c1", "c2"):
        assert torch.equal(gpu[k].cpu(), cpu[k]), k
