"""Pins oracle/vae_oracle.py against the reference's OWN VAE decoder and writes tests/golden/vae_decoder_outputs.pt.

Runs only in the build container (needs /root/reference):   python tests/golden/make_golden_vae.py
Imports the unmodified `external.sd1.ldm.modules.diffusionmodules.model.Decoder` / `.Encoder` (the reference's AutoencoderKL wrapper
additionally needs `taming`, which is not installed; its decode() is `decoder(post_quant_conv(z))`, autoencoder.py:331-334,
restated here with a plain nn.Conv2d; likewise encode() = quant_conv(encoder(x)), :325-329), loads the product's seeded state dict into it with strict=True — which proves the
parameter names and shapes of mvdfusion_b200.mvdfusion.autoencoder match the reference — and compares on seeded latents.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MVD_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shims"), REF, ROOT, os.path.join(ROOT, "tests")]

from common import rel_l2  # noqa: E402
from mvdfusion_b200 import synthetic  # noqa: E402
from mvdfusion_b200.mvdfusion.autoencoder import AutoencoderKL  # noqa: E402
from oracle import vae_oracle as V  # noqa: E402

from external.sd1.ldm.modules.diffusionmodules.model import Decoder as RefDecoder, Encoder as RefEncoder  # noqa: E402  (the reference)

SMALL = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 4, 4], num_res_blocks=2,
             attn_resolutions=[], dropout=0.0)
FULL = dict(SMALL, resolution=256, ch=128)  # configs/mvd_gso.yaml:53-71


def run(dd, n, seed):
    ours = AutoencoderKL(ddconfig=dd, lossconfig={"target": "torch.nn.Identity"}, embed_dim=4)
    synthetic.randomize_parameters(ours, seed)
    sd = {k: v.detach().float() for k, v in ours.state_dict().items()}
    ref = RefDecoder(**dd).eval()
    ref.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    pq = torch.nn.Conv2d(4, dd["z_channels"], 1)
    pq.load_state_dict({"weight": sd["post_quant_conv.weight"], "bias": sd["post_quant_conv.bias"]}, strict=True)
    S = dd["resolution"] // 8
    z = torch.randn(n, 4, S, S, generator=torch.Generator().manual_seed(seed + 1))
    with torch.no_grad():
        y_ref = ref(pq(z))
        y_orc = V.vae_decode(sd, z, ch_mult=dd["ch_mult"], num_res_blocks=dd["num_res_blocks"])
    r = rel_l2(y_orc, y_ref)
    print(f"vae decode ch={dd['ch']} res={dd['resolution']} n={n}: oracle vs reference rel-L2 = {r:.3e}, |y| max {y_ref.abs().max():.3f}")
    assert r < 2e-5, r
    # encode side: Encoder -> quant_conv -> moments (autoencoder.py:325-329); images in [-1, 1]
    enc = RefEncoder(**dd).eval()
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    qc = torch.nn.Conv2d(2 * dd["z_channels"], 8, 1)
    qc.load_state_dict({"weight": sd["quant_conv.weight"], "bias": sd["quant_conv.bias"]}, strict=True)
    img = torch.rand(n, 3, dd["resolution"], dd["resolution"], generator=torch.Generator().manual_seed(seed + 2)) * 2 - 1
    with torch.no_grad():
        m_ref = qc(enc(img))
        m_orc = V.vae_encode_moments(sd, img, ch_mult=dd["ch_mult"], num_res_blocks=dd["num_res_blocks"])
    r = rel_l2(m_orc, m_ref)
    print(f"vae encode ch={dd['ch']} res={dd['resolution']} n={n}: oracle vs reference rel-L2 = {r:.3e}, |moments| max {m_ref.abs().max():.3f}")
    assert r < 2e-5, r
    return {"ddconfig": dd, "seed": seed, "z": z, "y": y_ref, "img_seed": seed + 2, "moments": m_ref}


if __name__ == "__main__":
    out = {"small": run(SMALL, 2, 4321)}
    full = run(FULL, 1, 4322)
    # the full-size image is 3 x 256 x 256 floats: keep a strided sample of it (every 4th pixel) to stay small in git
    full["y_sample"] = full.pop("y")[:, :, ::4, ::4].clone()
    out["full"] = full
    torch.save(out, os.path.join(HERE, "vae_decoder_outputs.pt"))
    print("wrote", os.path.join(HERE, "vae_decoder_outputs.pt"))
