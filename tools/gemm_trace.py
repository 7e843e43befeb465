#!/usr/bin/env python
"""In-kernel timeline of every GEMM launch of one denoising step (B200 only; needs the instrumented library:
`make -C mvdfusion_b200/csrc trace`, run with MVD_B200_LIB=mvdfusion_b200/libmvd_b200_trace.so).

Each CTA of gemm_tc_kernel leaves a record (global timer at entry / exit, SM clock at the hand-over points of its warp roles);
the step is replayed from its CUDA graph, so the gaps are the real in-graph ones.  Prints, per GEMM shape: launch span, the gap
to the previous GEMM's end, and the median CTA's phases in microseconds.

    MVD_B200_LIB=mvdfusion_b200/libmvd_b200_trace.so python tools/gemm_trace.py [--views 8] [--out gpurun_out/gemm_trace.json]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

REC = np.dtype([("gt_entry", "<u8"), ("gt_exit", "<u8"), ("st", "<u4", (12,)), ("bid", "<i4"), ("grid", "<i4"), ("M", "<i4"), ("N", "<i4"),
                ("num_kb", "<i4"), ("BN", "<i4"), ("split", "<i4"), ("n_local", "<i4"), ("flags", "<i4"), ("pad", "<i4")])
NAMES = ["entry", "setup", "prod_wait", "prod_unit0", "mma_first", "mma_unit0", "mma_last", "epi_wait", "epi_first", "epi_lastacc", "epi_done", "exit"]


def read_trace(lib, cap=1 << 17):
    buf = np.zeros(cap, dtype=REC)
    n = lib.mvd_debug_gemm_trace(ctypes.c_void_p(buf.ctypes.data), ctypes.c_int(cap))
    if n < 0:
        raise RuntimeError("mvd_debug_gemm_trace failed")
    return buf[:n]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--cfg", type=float, default=2.5)
    ap.add_argument("--out", default="")
    ap.add_argument("--mhz", type=float, default=1965.0)
    a = ap.parse_args()
    from common import build_model, synthetic
    from mvdfusion_b200 import _lib
    from mvdfusion_b200.mvdfusion.cameras import PerspectiveCameras
    from mvdfusion_b200.runtime import current_stream
    lib = _lib.load()
    if not hasattr(lib, "mvd_debug_gemm_trace"):
        raise SystemExit("the loaded library has no trace entry point: set MVD_B200_LIB to libmvd_b200_trace.so (make -C mvdfusion_b200/csrc trace)")
    lib.mvd_debug_gemm_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.mvd_debug_gemm_trace.restype = ctypes.c_int
    dev = torch.device("cuda", 0)
    n, S, D = a.views, 32, 1
    model = build_model(320, 8, D=D, S=S, device=dev)
    sc = synthetic.scene_inputs(n, S)
    de, dn = synthetic.step_noises(n, D, S, 8, seed=1)
    rows = torch.stack([model.ddim.step_row(49 - i, a.cfg) for i in range(8)])
    cam = lambda c: PerspectiveCameras(c["R"], c["T"], c["f"], c["p"], device=dev)
    plan = model.step_plan(n, S, D, use_cfg=a.cfg != 1.0)
    stream = current_stream(dev)
    model.bind_scene(plan, cam(sc["cams"]), sc["input_latents"].to(dev), cam(sc["in_cams"]), sc["clip_v_embed"].to(dev), stream)
    plan.x.copy_(sc["x_T"].reshape(n, 5, S * S))
    plan.set_tables(rows, de, dn)
    for _ in range(3):
        plan.loop_step(stream)  # capture + warm
    torch.cuda.synchronize()
    read_trace(lib)  # drop
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.loop_step(stream)
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1)
    recs = read_trace(lib)
    print(f"one graph step: {step_ms:.3f} ms, {len(recs)} CTA records")
    # split into launches: records of one launch share (grid, M, N, num_kb, BN, split, flags) and overlap in time
    order = np.argsort(recs["gt_entry"], kind="stable")
    recs = recs[order]
    launches, cur, key = [], [], None
    for r in recs:
        k = (int(r["grid"]), int(r["M"]), int(r["N"]), int(r["num_kb"]), int(r["BN"]), int(r["split"]), int(r["flags"]))
        if key is not None and (k != key or len(cur) >= key[0]):
            launches.append((key, cur))
            cur = []
        key = k
        cur.append(r)
    if cur:
        launches.append((key, cur))
    t_step0 = min(int(r["gt_entry"]) for r in recs)
    cyc = 1e3 / a.mhz  # ns per cycle
    out, prev_end = [], None
    for key, rs in launches:
        rs = np.array(rs, dtype=REC)
        t0, t1 = int(rs["gt_entry"].min()), int(rs["gt_exit"].max())
        st = rs["st"].astype(np.int64)
        rel = ((st - st[:, :1]) & 0xFFFFFFFF).astype(np.float64) * cyc / 1e3  # us since CTA entry
        rel[st == 0] = np.nan
        med = np.nanmedian(rel, axis=0)
        late_entry = (rs["gt_entry"].max() - t0) / 1e3
        out.append({"grid": key[0], "M": key[1], "N": key[2], "num_kb": key[3], "BN": key[4], "split": key[5], "pair": key[6] & 1, "res": (key[6] >> 1) & 1,
                    "act": (key[6] >> 4) & 15, "out_mode": (key[6] >> 8) & 15, "conv": (key[6] >> 12) & 15, "nwg": key[6] >> 16, "ctas": len(rs),
                    "start_us": (t0 - t_step0) / 1e3, "span_us": (t1 - t0) / 1e3, "gap_prev_us": None if prev_end is None else (t0 - prev_end) / 1e3,
                    "entry_spread_us": late_entry, "n_local_max": int(rs["n_local"].max()), "phases_us": {n_: (None if np.isnan(v) else round(float(v), 2)) for n_, v in zip(NAMES, med)}})
        prev_end = t1
    print(f"{len(out)} GEMM launches; sum of spans {sum(o['span_us'] for o in out) / 1e3:.3f} ms")
    groups = {}
    for o in out:
        k = (o["conv"], o["M"], o["N"], o["num_kb"], o["BN"], o["split"], o["pair"], o["res"], o["act"], o["out_mode"])
        groups.setdefault(k, []).append(o)
    print("conv M N kb BN split pair res act out | calls  span  gap_prev | " + " ".join(f"{n_[:9]:>9s}" for n_ in NAMES[1:]))
    for k, os_ in sorted(groups.items(), key=lambda kv: -sum(o["span_us"] for o in kv[1])):
        span = np.median([o["span_us"] for o in os_])
        gap = np.median([o["gap_prev_us"] for o in os_ if o["gap_prev_us"] is not None] or [0])
        ph = [np.nanmedian([o["phases_us"][n_] if o["phases_us"][n_] is not None else np.nan for o in os_]) for n_ in NAMES[1:]]
        print(" ".join(str(x) for x in k) + f" | x{len(os_):2d} {span:6.1f} {gap:6.1f} | " + " ".join(f"{v:9.2f}" for v in ph))
    if a.out:
        json.dump({"step_ms": step_ms, "launches": out}, open(a.out, "w"))


if __name__ == "__main__":
    main()
