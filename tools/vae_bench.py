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
    ap.add_argument("--table", action="store_true", help="per-kernel device times of one decode (CUDA events, eager)")
    ap.add_argument("--cpu", action="store_true", help="also time the reference algorithm (oracle/vae_oracle.py, fp32) on the host cores, 1 view")
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
    line = {"what": "AutoencoderKL.decode", "views": a.views, "ms": round(best, 3), "launches": launches,
            "tflops": round(gflop / best, 1), "finite": bool(torch.isfinite(y).all())}
    # encode: the scene's input image(s) before the loop (viewfusion_zero_depth_rgb.py:158-159); 273 GFLOP per 256^2 image
    img = torch.rand(a.views, 3, 256, 256, device="cuda") * 2 - 1
    post = m.encode(img)
    torch.cuda.synchronize()
    best_e = 1e30
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        post = m.encode(img)
        e1.record()
        torch.cuda.synchronize()
        best_e = min(best_e, e0.elapsed_time(e1))
    line["encode"] = {"views": a.views, "ms": round(best_e, 3), "tflops": round(273.0 * a.views / best_e, 1),
                      "finite": bool(torch.isfinite(post.parameters).all())}
    if a.table:
        plan = m.__dict__["_mvd_cache"].plans[("decode", a.views, 32)]
        calls = plan.prog.calls
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in calls]
        st = torch.cuda.current_stream().cuda_stream
        torch.cuda.synchronize()
        torch.cuda._sleep(int(1.0e8))
        for c, (e0, e1) in zip(calls, ev):
            e0.record()
            c(st)
            e1.record()
        torch.cuda.synchronize()
        agg = {}
        for c, (e0, e1) in zip(calls, ev):
            key = (c.meta.get("kernel", c.name.replace("mvd_", "")), str(c.meta.get("desc", "")))
            g = agg.setdefault(key, [0, 0.0, 0.0])
            g[0] += 1
            g[1] += e0.elapsed_time(e1)
            g[2] += c.meta.get("flops", 0.0)
        for (k, d), (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
            print(f"{k:22s} {d:48s} x{cnt:3d} {ms * 1e3 / cnt:9.1f} us/call {ms:8.3f} ms  {fl / (ms * 1e-3) / 1e12 if fl else 0:7.1f} TF/s", file=sys.stderr)
    if a.cpu:
        import time
        sys.path.insert(0, ROOT)
        from oracle import vae_oracle as V
        sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
        torch.set_num_threads(os.cpu_count())
        z1 = z[:1].cpu()
        with torch.no_grad():
            V.vae_decode(sd, z1[:, :, :8, :8])  # warm-up on a small map
            t0 = time.perf_counter()
            V.vae_decode(sd, z1)
            dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"ms_per_view": round(dt * 1e3, 1), "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "1 view, fp32, oracle/vae_oracle.py", "ms_for_these_views": round(dt * 1e3 * a.views, 1)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
