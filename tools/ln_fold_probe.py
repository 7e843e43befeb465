#!/usr/bin/env python
"""The QKV / GEGLU GEMMs of one transformer block at the 32^2 level (M = 16384, C = 320), plain (LayerNorm pass in front) and with the
LayerNorm folded in (ABI 13), launched in the order [qkv, qkv+ln, geglu, geglu+ln] `--rounds` times: the subject of an ncu capture
(`ncu -k regex:gemm_tc_kernel -s 4 -c 4 ...`) and, without ncu, a graph-timed A/B of the four.

    python tools/ln_fold_probe.py [--M 16384] [--C 320] [--rounds 2] [--time]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=16384)
    ap.add_argument("--C", type=int, default=320)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    from mvdfusion_b200 import ops as OPS
    from ops_double import geglu_permutation
    ops = OPS.NativeOps("cuda:0")
    dev = "cuda:0"
    M, C, heads = a.M, a.C, 8
    d = C // heads
    dpad = (d + 63) // 64 * 64
    seq = 1024 if M % 1024 == 0 else M
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(M, C, generator=g) * 2.0
    ch = x.reshape(M, C // 32, 32)
    st = torch.stack([ch.sum(-1), (ch * ch).sum(-1)], dim=-1).permute(1, 0, 2).contiguous().to(dev)
    x16 = x.half().to(dev)
    ln16 = torch.nn.functional.layer_norm(x, (C,)).half().to(dev)
    Wq = (torch.randn(3 * C, C, generator=g) * C ** -0.5).half().to(dev)
    Wg = (torch.randn(8 * C, C, generator=g) * C ** -0.5)[geglu_permutation(4 * C, 256)].half().to(dev)
    csq, csg = Wq.float().sum(1), Wg.float().sum(1)
    bq, bg = torch.randn(3 * C, generator=g).to(dev), torch.randn(8 * C, generator=g).to(dev)
    nb = (M // seq) * heads * seq * dpad
    q, k, vt = (torch.zeros(nb, dtype=torch.float16, device=dev) for _ in range(3))
    o = torch.zeros(M, 4 * C, dtype=torch.float16, device=dev)
    qkv = dict(out_k=k, out_vt=vt, heads=heads, dhead=d, dpad=dpad, seq=seq)
    calls = [
        ("qkv", ops.gemm(ln16, Wq, q, M, 3 * C, C, qkv=qkv)),
        ("qkv+ln", ops.gemm(x16, Wq, q, M, 3 * C, C, qkv=qkv, bias=bq, ln=(st, csq, 1e-5))),
        ("geglu", ops.gemm(ln16, Wg, o, M, 8 * C, C, bias=bg, act=OPS.ACT_GEGLU, tile_n=256, ldc=4 * C)),
        ("geglu+ln", ops.gemm(x16, Wg, o, M, 8 * C, C, bias=bg, act=OPS.ACT_GEGLU, tile_n=256, ldc=4 * C, ln=(st, csg, 1e-5))),
    ]
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(a.rounds):
        for _, c in calls:
            c(s)
    torch.cuda.synchronize()
    if a.time:
        for name, c in calls:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(20):
                    c(torch.cuda.current_stream().cuda_stream)
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(3):
                e0.record()
                gr.replay()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / 20)
            print(f"{name:10s} {best:7.2f} us", flush=True)


if __name__ == "__main__":
    main()
