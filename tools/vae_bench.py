#!/usr/bin/env python
"""Device time of the VAE decode that follows the denoising loop: n views of 32x32x4 latents ->
n x 3 x 256 x 256 images, full-size decoder (configs/mvd_gso.yaml:53-71), random-init weights.  B200 only.

    python tools/vae_bench.py [--views 8] [--reps 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from mvdfusion_b200 import _lib, synthetic
    from mvdfusion_b200.mvdfusion.autoencoder import AutoencoderKL
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2,
              attn_resolutions=[], dropout=0.0)
    m = AutoencoderKL(ddconfig=dd, embed_dim=4)
    synthetic.randomize_parameters(m, 7)
    m = m.cuda().eval()
    z = torch.randn(a.views, 4, 32, 32, device="cuda")
    n0 = _lib.launch_count()
    y = m.decode(z)
    torch.cuda.synchronize()
    launches = _lib.launch_count() - n0
    best = 1e30
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = m.decode(z)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    gflop = 622.0 * a.views  # 622 GFLOP per 256^2 decode
    print(json.dumps({"what": "AutoencoderKL.decode", "views": a.views, "ms": round(best, 3), "launches": launches,
                      "tflops": round(gflop / best, 1), "finite": bool(torch.isfinite(y).all())}))


if __name__ == "__main__":
    main()
