"""VAE decode / encode: oracle vs the reference golden, product host logic vs oracle (CPU emulation of the C ABI),
and — on the B200 — the kernels themselves."""
import os

import pytest
import torch

from common import rel_l2
from mvdfusion_b200 import synthetic
from oracle import vae_oracle as V

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vae_decoder_outputs.pt")
# fp16 tensor-core operands, fp32 accumulation / residual stream / norms, as on the denoising path.  The decoder stacks ~60
# convolutions without the UNet's zero-initialised-branch structure and ends in an 8-bit image (1/255 = 3.9e-3), so its gate
# is 5e-3 rel-L2; measured: 3.2e-3 for the 32-channel test decoder, see profiles/ for the full-size figure on the B200.
TOL = 5e-3


def build_vae(dd, seed, device="cpu"):
    from mvdfusion_b200.mvdfusion.autoencoder import AutoencoderKL
    m = AutoencoderKL(ddconfig=dd, lossconfig={"target": "torch.nn.Identity"}, embed_dim=4)
    synthetic.randomize_parameters(m, seed)
    return m.to(device).eval()


def sd_of(m):
    return {k: v.detach().float().cpu() for k, v in m.state_dict().items()}


def test_oracle_matches_the_reference_decoder_goldens():
    g = torch.load(GOLD)
    for name in ("small", "full"):
        e = g[name]
        dd = e["ddconfig"]
        if name == "full" and os.environ.get("MVD_FAST_TESTS"):
            continue
        m = build_vae(dd, e["seed"])
        with torch.no_grad():
            y = V.vae_decode(sd_of(m), e["z"], ch_mult=dd["ch_mult"], num_res_blocks=dd["num_res_blocks"])
        ref = e["y"] if "y" in e else e["y_sample"]
        y = y if "y" in e else y[:, :, ::4, ::4]
        assert rel_l2(y, ref) < 2e-5, name


def golden_image(e):
    dd = e["ddconfig"]
    n = e["z"].shape[0]
    return torch.rand(n, 3, dd["resolution"], dd["resolution"], generator=torch.Generator().manual_seed(e["img_seed"])) * 2 - 1


def test_oracle_matches_the_reference_encoder_goldens():
    g = torch.load(GOLD)
    for name in ("small", "full"):
        e = g[name]
        dd = e["ddconfig"]
        if name == "full" and os.environ.get("MVD_FAST_TESTS"):
            continue
        m = build_vae(dd, e["seed"])
        with torch.no_grad():
            mom = V.vae_encode_moments(sd_of(m), golden_image(e), ch_mult=dd["ch_mult"], num_res_blocks=dd["num_res_blocks"])
        assert rel_l2(mom, e["moments"]) < 2e-5, name


def test_encode_host_logic_vs_oracle(ops_double):
    e = torch.load(GOLD)["small"]
    m = build_vae(e["ddconfig"], e["seed"])
    post = m.encode(golden_image(e))
    assert post.parameters.shape == e["moments"].shape and torch.isfinite(post.parameters).all()
    assert rel_l2(post.parameters, e["moments"]) < TOL
    assert torch.equal(post.mode(), post.parameters[:, :4]) and post.sample().shape == post.mean.shape


def test_yaml_target_resolves_to_this_package_vae():
    from mvdfusion_b200.config import instantiate_from_config
    from mvdfusion_b200.mvdfusion.autoencoder import AutoencoderKL
    dd = torch.load(GOLD)["small"]["ddconfig"]
    m = instantiate_from_config({"target": "external.sd1.ldm.models.autoencoder.AutoencoderKL",
                                 "params": {"embed_dim": 4, "monitor": "val/rec_loss", "ddconfig": dd,
                                            "lossconfig": {"target": "torch.nn.Identity"}}})
    assert isinstance(m, AutoencoderKL)
    keys = m.state_dict().keys()
    for k in ("post_quant_conv.weight", "decoder.conv_in.bias", "decoder.mid.attn_1.proj_out.weight", "decoder.up.3.upsample.conv.weight",
              "decoder.up.1.block.0.nin_shortcut.weight", "decoder.up.0.block.2.conv2.bias", "decoder.norm_out.weight", "decoder.conv_out.weight"):
        assert k in keys, k
    assert "encoder.down.2.downsample.conv.weight" in keys and "quant_conv.bias" in keys


def test_decode_host_logic_vs_oracle(ops_double):
    e = torch.load(GOLD)["small"]
    dd = e["ddconfig"]
    m = build_vae(dd, e["seed"])
    y = m.decode(e["z"])
    assert y.shape == e["y"].shape and torch.isfinite(y).all()
    assert rel_l2(y, e["y"]) < TOL
    # ViewFusion.decode's tail (viewfusion_zero_depth_rgb.py:162-163)
    with torch.no_grad():
        img = V.viewfusion_decode(sd_of(m), e["z"] * 0.18215, ch_mult=dd["ch_mult"], num_res_blocks=dd["num_res_blocks"])
    assert rel_l2(((y + 1.0) / 2.0).clip(0.0, 1.0), img) < TOL


def test_viewfusion_decode_through_the_facade(ops_double):
    """ViewFusion(vae_config=...) builds the VAE from the reference's yaml section; .decode(z) = unnormalize(vae.decode(z / 0.18215))"""
    from common import model_config
    from mvdfusion_b200.config import instantiate_from_config
    e = torch.load(GOLD)["small"]
    dd = e["ddconfig"]
    cfg = model_config(64, 8, D=1, S=8)
    cfg["params"]["vae_config"] = {"target": "external.sd1.ldm.models.autoencoder.AutoencoderKL",
                                   "params": {"embed_dim": 4, "monitor": "val/rec_loss", "ddconfig": dd,
                                              "lossconfig": {"target": "torch.nn.Identity"}}}
    m = instantiate_from_config(cfg).eval()
    assert m.vae is not None
    synthetic.randomize_parameters(m.vae, e["seed"])
    z = e["z"] * 0.18215
    img = m.decode(z)
    with torch.no_grad():
        ref = V.viewfusion_decode(sd_of(m.vae), z, ch_mult=dd["ch_mult"], num_res_blocks=dd["num_res_blocks"])
    assert img.shape == ref.shape and float(img.min()) >= 0.0 and float(img.max()) <= 1.0
    assert rel_l2(img, ref) < TOL
    assert any(k.startswith("vae.decoder.mid.attn_1.") for k in m.state_dict())
    # ViewFusion.encode (viewfusion_zero_depth_rgb.py:158-159): images in [0, 1] -> scaled latents
    pic = (golden_image(e) + 1) / 2
    lat = m.encode(pic)
    with torch.no_grad():
        lat_ref = V.viewfusion_encode(sd_of(m.vae), pic, ch_mult=dd["ch_mult"], num_res_blocks=dd["num_res_blocks"])
    assert lat.shape == lat_ref.shape and rel_l2(lat, lat_ref) < TOL


@pytest.mark.gpu
def test_softmax_rows_and_wide_convolution_kernels():
    from mvdfusion_b200 import ops as OPS
    from ops_double import TorchOpsDouble
    nat, dbl = OPS.NativeOps("cuda:0"), TorchOpsDouble()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(5)
    for rows, cols in ((1024, 1024), (64, 64), (300, 520)):
        s = torch.randn(rows, cols, generator=g) * 3
        p_cpu, p_gpu = torch.zeros(rows, cols, dtype=torch.float16), torch.zeros(rows, cols, dtype=torch.float16, device="cuda")
        dbl.softmax_rows(s, p_cpu, rows, cols, 0.05)(None)
        nat.softmax_rows(s.cuda(), p_gpu, rows, cols, 0.05)(st)
        torch.cuda.synchronize()
        assert (p_gpu.cpu().float() - p_cpu.float()).abs().max() < 2e-3 * p_cpu.float().abs().max()
    # conv3x3 on maps wider than one 128-pixel tile (the decoder's 256-wide levels), odd output width included
    for n, H, Wd, Cin, Cout in ((1, 256, 256, 32, 32), (2, 128, 256, 64, 3)):
        M = n * H * Wd
        ldc = 4 if Cout == 3 else Cout
        A = (torch.randn(M, Cin, generator=g)).half()
        Wt = (torch.randn(Cout, 9 * Cin, generator=g) * (9 * Cin) ** -0.5).half()
        bias = torch.randn(Cout, generator=g)
        o_cpu, o_gpu = torch.zeros(M, ldc), torch.zeros(M, ldc, device="cuda")
        dbl.gemm(A, Wt, o_cpu, M, Cout, 9 * Cin, conv=(n, H, Wd, Cin), bias=bias, ldc=ldc)(None)
        nat.gemm(A.cuda(), Wt.cuda(), o_gpu, M, Cout, 9 * Cin, conv=(n, H, Wd, Cin), bias=bias.cuda(), ldc=ldc)(st)
        torch.cuda.synchronize()
        assert (o_gpu.cpu() - o_cpu).abs().max() < 2e-3 * o_cpu.abs().max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small", "full"])
def test_encode_on_the_gpu_vs_reference_golden(name):
    from common import record_parity
    e = torch.load(GOLD)[name]
    m = build_vae(e["ddconfig"], e["seed"], device="cuda")
    post = m.encode(golden_image(e).cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(post.parameters).all()
    assert record_parity(f"vae_encode_{name}_vs_reference_golden", rel_l2(post.parameters, e["moments"]), TOL) < TOL


@pytest.mark.gpu
def test_stride2_im2col_without_low_padding():
    from mvdfusion_b200 import ops as OPS
    from ops_double import TorchOpsDouble
    nat, dbl = OPS.NativeOps("cuda:0"), TorchOpsDouble()
    n, H, C = 2, 16, 32
    x = torch.randn(n * H * H, C, generator=torch.Generator().manual_seed(3))
    for pad_lo in (0, 1):
        y_cpu = torch.zeros(n * (H // 2) ** 2, 9 * C, dtype=torch.float16)
        y_gpu = torch.zeros_like(y_cpu, device="cuda")
        dbl.im2col_s2(x, y_cpu, n, H, H, C, pad_lo=pad_lo)(None)
        nat.im2col_s2(x.cuda(), y_gpu, n, H, H, C, pad_lo=pad_lo)(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(y_gpu.cpu(), y_cpu)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small", "full"])
def test_decode_on_the_gpu_vs_reference_golden(name):
    from common import record_parity
    e = torch.load(GOLD)[name]
    m = build_vae(e["ddconfig"], e["seed"], device="cuda")
    y = m.decode(e["z"].cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    ref = e["y"] if "y" in e else e["y_sample"]
    y = y if "y" in e else y[:, :, ::4, ::4]
    assert record_parity(f"vae_decode_{name}_vs_reference_golden", rel_l2(y, ref), TOL) < TOL
