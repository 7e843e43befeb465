#!/usr/bin/env python
"""Per-launch device times of one denoising step, grouped by kernel and shape (CUDA events; B200 only).

    python tools/step_profile.py [--views 8] [--reps 5] [--out gpurun_out/step_profile.json]
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
    ap.add_argument("--cfg", type=float, default=2.5)
    ap.add_argument("--out", default="")
    ap.add_argument("--world", type=int, default=1, help="profile rank 0's step of a scene view-sharded over this many GPUs (no collective; one GPU suffices)")
    a = ap.parse_args()
    from common import build_model, synthetic
    from mvdfusion_b200.mvdfusion.cameras import PerspectiveCameras
    from mvdfusion_b200.runtime import current_stream
    dev = torch.device("cuda", 0)
    n, S, D = a.views, 32, 1
    model = build_model(320, 8, D=D, S=S, device=dev)
    sc = synthetic.scene_inputs(n, S)
    de, dn = synthetic.step_noises(n, D, S, 4, seed=1)
    rows = torch.stack([model.ddim.step_row(49 - i, a.cfg) for i in range(4)])
    cam = lambda c: PerspectiveCameras(c["R"], c["T"], c["f"], c["p"], device=dev)
    if a.world > 1:
        model.view_group = (None, 0, a.world)  # rank 0's shard: the shapes are the same on every rank
    plan = model.step_plan(n, S, D, use_cfg=a.cfg != 1.0)
    stream = current_stream(dev)
    model.bind_scene(plan, cam(sc["cams"]), sc["input_latents"].to(dev), cam(sc["in_cams"]), sc["clip_v_embed"].to(dev), stream)
    plan.x.copy_(sc["x_T"].reshape(n, 5, S * S))
    plan.set_tables(rows, de, dn)
    calls = [c for c in plan._loop_prog.calls if getattr(c, "name", "") != "all_gather_latents"]
    # the whole step from its CUDA graph (what the sampler replays), without the collective
    if a.world == 1:
        for _ in range(3):
            plan.loop_step(stream)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        plan.counter.zero_()
        g0.record()
        for _ in range(4):
            plan.loop_step(stream)
        g1.record()
        torch.cuda.synchronize()
        print(f"graph replay: {g0.elapsed_time(g1) / 4:.3f} ms per step")
    else:
        gr = torch.cuda.CUDAGraph()
        for c in calls:
            c(stream)
        torch.cuda.synchronize()
        with torch.cuda.graph(gr):
            for c in calls:
                c(torch.cuda.current_stream().cuda_stream)
        gr.replay()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        plan.counter.zero_()
        g0.record()
        for _ in range(4):
            gr.replay()
        g1.record()
        torch.cuda.synchronize()
        print(f"graph replay of rank 0's {len(calls)} kernels (1/{a.world} of the views, no all-gather): {g0.elapsed_time(g1) / 4:.3f} ms per step")
    acc = [0.0] * len(calls)
    for rep in range(a.reps + 1):
        plan.counter.zero_()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in calls]
        torch.cuda.synchronize()
        torch.cuda._sleep(int(1.5e8))
        for c, (e0, e1) in zip(calls, ev):
            e0.record()
            c(stream)
            e1.record()
        torch.cuda.synchronize()
        if rep:
            for i, (e0, e1) in enumerate(ev):
                acc[i] += e0.elapsed_time(e1) / a.reps
    groups = {}
    for c, ms in zip(calls, acc):
        name = c.meta.get("kernel", c.name.replace("mvd_", ""))
        key = (name, str(c.meta.get("desc", c.meta.get("shape", ""))))
        g = groups.setdefault(key, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        g["calls"] += 1
        g["ms"] += ms
        g["flops"] += c.meta.get("flops", 0.0)
        g["bytes"] += c.meta.get("bytes", 0.0)
    total = sum(acc)
    print(f"step: {len(calls)} calls, sum of kernel times {total:.3f} ms")
    out = []
    for (name, shape), g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
        tf = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["flops"] else 0.0
        gbs = g["bytes"] / (g["ms"] * 1e-3) / 1e9 if g["bytes"] else 0.0
        print(f"{name:24s} {shape:60s} x{g['calls']:3d} {g['ms']:8.3f} ms {100 * g['ms'] / total:5.1f}%  {g['ms'] / g['calls'] * 1e3:8.1f} us/call "
              f"{tf:7.1f} TF/s {gbs:7.0f} GB/s")
        out.append({"kernel": name, "shape": shape, **g})
    if a.out:
        json.dump({"total_ms": total, "groups": out}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
