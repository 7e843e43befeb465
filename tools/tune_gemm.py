#!/usr/bin/env python
"""Measure (tile_n, split_k, cta_pair) for every distinct GEMM launch of the denoising step on this GPU and write
mvdfusion_b200/gemm_tuning.json (engine.Builder.gemm consults it; without the file the library's heuristics decide).

    python tools/tune_gemm.py [--views 8] [--out mvdfusion_b200/gemm_tuning.json]

Each candidate is timed as a CUDA graph of `reps` launches that rotate over copies of the weight matrix (> L2 in total),
so weights stream from HBM as in the real step while the activations stay L2-resident.
"""
import argparse
import copy
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def time_candidate(lib, g0, Wt, copies, tile_n, split_k, pair, reps, stream):
    gs = []
    for i in range(reps):
        g = type(g0)()
        ctypes.memmove(ctypes.byref(g), ctypes.byref(g0), ctypes.sizeof(g0))
        g.Wt = copies[i % len(copies)].data_ptr()
        g.tile_n, g.split_k, g.cta_pair = tile_n, split_k, pair
        gs.append(g)
    rc = lib.mvd_gemm_f16(ctypes.byref(gs[0]), stream.cuda_stream)
    if rc != 0:
        return None
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(graph, stream=stream):
            for g in gs:
                if lib.mvd_gemm_f16(ctypes.byref(g), torch.cuda.current_stream().cuda_stream) != 0:
                    return None
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(2):
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--reps", type=int, default=16)
    ap.add_argument("--latent", type=int, default=32, help="latent side (32 = 256^2 images, 64 = 512^2)")
    ap.add_argument("--shards", default="1", help="comma list of GPU counts whose per-rank shapes (views/G per GPU) to tune")
    ap.add_argument("--merge", default="", help="existing tuning file whose choices are kept for signatures not re-measured")
    ap.add_argument("--only", default="", help="re-measure only the signatures that contain this substring (e.g. '+st')")
    ap.add_argument("--out", default=os.path.join(ROOT, "mvdfusion_b200", "gemm_tuning.json"))
    a = ap.parse_args()
    os.environ["MVD_GEMM_NO_TUNING"] = "1"  # start from the heuristics
    from common import build_model
    from mvdfusion_b200 import _lib
    from mvdfusion_b200.ops import ACT_GEGLU
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    model = build_model(320, 8, D=1, S=a.latent, device=dev)
    stream = torch.cuda.Stream()
    seen, choices, report = {}, {}, []
    if a.merge and os.path.exists(a.merge):
        choices.update(json.load(open(a.merge)).get("choices", {}))
    for world in [int(w) for w in a.shards.split(",")]:
        model.view_group = (None, 0, world) if world > 1 else None  # rank 0's shard: shapes are the same on every rank
        plan = model.step_plan(a.views, a.latent, 1, use_cfg=True)
        print(f"---- {a.views} views over {world} GPU(s): {plan.q} views per GPU", flush=True)
        for c in plan.core_prog.calls:
            if c.name != "mvd_gemm_f16":
                continue
            sig = c.meta["sig"]
            if sig in seen or (a.only and a.only not in sig):
                continue
            seen[sig] = True
            choices.pop(sig, None)
            g0, Wt = c.keep[0], c.keep[2]
            M, N, K = g0.M, g0.N, g0.K
            wbytes = Wt.numel() * 2
            ncopy = max(1, min(24, (256 << 20) // wbytes + 1))
            copies = [Wt] + [Wt.clone() for _ in range(ncopy - 1)]
            can_split = bool(c.meta["can_split"])
            geglu = g0.act == ACT_GEGLU
            tiles_m = (M + 127) // 128
            base = time_candidate(lib, g0, Wt, copies, g0.tile_n, g0.split_k, 0, a.reps, stream)
            cands = []
            for pair in (1, 2):
                if pair == 2 and tiles_m < 2:
                    continue
                # narrow tiles for the one / two m-tile layers: more CTAs stream the weight matrix in parallel
                for tn in ((g0.tile_n,) if geglu else ((32, 48) if (tiles_m <= 2 and pair == 1) else ()) + (64, 96, 128, 160, 192, 224, 256)):
                    if not geglu and N <= 64:
                        tn = 0
                    for sk in ((1, 2, 3, 4, 6, 8, 12, 16) if can_split else (1,)):
                        cands.append((tn, sk, pair))
            if not geglu and N % 320 == 0 and tiles_m >= 2:
                cands.append((320, 1, 2))  # one 320-column pair tile per two m-tiles (two N = 160 MMAs, single accumulator, TMA epilogue)
            cands = sorted(set(cands))
            best, best_c = base, None
            for tn, sk, pair in cands:
                t = time_candidate(lib, g0, Wt, copies, tn, sk, pair, a.reps, stream)
                if t is not None and t < best * 0.97:  # only move off the heuristic for a clear win
                    best, best_c = t, (tn, sk, pair)
            if best_c is not None:
                choices[sig] = list(best_c)
            report.append((sig, base, best, best_c))
            print(f"{sig:44s} heuristic {base:7.1f} us   best {best:7.1f} us  {best_c}", flush=True)
            del copies
    gain = sum((b - t) * 1 for _, b, t, _ in report)
    json.dump({"device": torch.cuda.get_device_name(0), "views": a.views, "note": "signature -> [tile_n, split_k, cta_pair]",
               "choices": choices}, open(a.out, "w"), indent=1, sort_keys=True)
    print(f"{len(choices)} of {len(report)} shapes tuned; summed per-shape gain {gain:.1f} us (x call counts in the step)")


if __name__ == "__main__":
    main()
