"""Roofline of the training-side streaming kernels (csrc/train.cu, ABI 15) on one B200: each entry point at the shapes of a
BASELINE configs[3] training step (N = 8 views, 256^2, D = 3), timed with CUDA events on the launching stream over buffers that
rotate through more than the 126 MB L2 (so every pass streams from HBM), against the measured copy bandwidth of MEASURED_PEAKS.json.

    python tools/train_kernels_bench.py > profiles/r02_train_kernels_bench.json

`algo_bytes` = the bytes the operation must move once (inputs read once + outputs written once); `moved_bytes` = what the kernels
actually touch (LayerNorm backward reads dy and x a second time for dgamma / dbeta, GroupNorm reads x twice forward and (dy, x)
twice backward) — the second reads are what a fused single-pass variant would save.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvdfusion_b200 import ops as OPS  # noqa: E402


def timed(calls, iters=20, warm=3):
    st = torch.cuda.current_stream().cuda_stream
    for i in range(warm):
        calls[i % len(calls)](st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        calls[i % len(calls)](st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6548.8))
    ops = OPS.NativeOps("cuda:0")
    r = lambda *s: torch.randn(*s, device="cuda")
    rows = []

    def copies(elems_bytes):
        return max(2, int(300e6 // elems_bytes) + 1)

    def record(name, us, algo, moved):
        rows.append({"kernel": name, "us": round(us, 2), "algo_bytes": algo, "moved_bytes": moved, "algo_gbs": round(algo / us * 1e-3, 1),
                     "moved_gbs": round(moved / us * 1e-3, 1), "frac_of_hbm_peak_algo": round(algo / us * 1e-3 / peak, 3),
                     "frac_of_hbm_peak_moved": round(moved / us * 1e-3 / peak, 3)})
        print(rows[-1], file=sys.stderr, flush=True)

    # LayerNorm: transformer blocks at 32^2 (8192 x 320) and the DiT blocks of GridAttn (N^2 HW D = 196608 rows x 256)
    for R, C, affine in ((8192, 320, True), (196608, 256, True)):
        k = copies(R * C * 4)
        xs, ys, ds = [r(R, C) for _ in range(k)], [torch.empty(R, C, device="cuda") for _ in range(k)], [r(R, C) for _ in range(k)]
        g, b, st = r(C), r(C), torch.empty(R, 2, device="cuda")
        dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
        us = timed([ops.layernorm_fwd(xs[i], g, b, ys[i], st, R, C, 1e-5) for i in range(k)])
        record(f"layernorm_fwd {R}x{C}", us, 8 * R * C, 8 * R * C)
        us = timed([ops.layernorm_bwd(ds[i], xs[i], g, st, ys[i], dg, db, R, C) for i in range(k)])
        record(f"layernorm_bwd {R}x{C} (+dgamma, dbeta)", us, 12 * R * C, 20 * R * C)
        us = timed([ops.layernorm_bwd(ds[i], xs[i], g, st, ys[i], None, None, R, C) for i in range(k)])
        record(f"layernorm_bwd {R}x{C} (dx only)", us, 12 * R * C, 12 * R * C)
        del xs, ys, ds
    # GroupNorm32 + SiLU: ResBlock in_layers at 32^2 x 320 and 16^2 x 640, the widest skip concatenation (32^2 x 960)
    for n, hw, C in ((8, 1024, 320), (8, 256, 640), (8, 1024, 960)):
        e = n * hw * C
        k = copies(e * 4)
        xs, ys, ds = [r(n, hw, C) for _ in range(k)], [torch.empty(n, hw, C, device="cuda") for _ in range(k)], [r(n, hw, C) for _ in range(k)]
        g, b = r(C), r(C)
        st, ws = torch.empty(n, 32, 2, device="cuda"), torch.empty(n * 16384, dtype=torch.uint8, device="cuda")
        dgb = torch.empty(2, C, device="cuda")
        dg, db = dgb[0], dgb[1]
        us = timed([ops.groupnorm_fwd(xs[i], g, b, ys[i], st, ws, n, hw, C, 1e-5, True) for i in range(k)])
        record(f"groupnorm_silu_fwd {n}x{hw}x{C}", us, 8 * e, 12 * e)
        us = timed([ops.groupnorm_bwd(ds[i], xs[i], g, b, st, ys[i], dg, db, ws, n, hw, C, True) for i in range(k)])
        record(f"groupnorm_silu_bwd {n}x{hw}x{C}", us, 12 * e, 20 * e)
        del xs, ys, ds
    # GEGLU of the feed-forward at 32^2 (8192 rows, inner 1280) and GELU of the DiT mlp (196608 x 512)
    R, I = 8192, 1280
    k = copies(R * 2 * I * 4)
    hs, ys, ds, dh = [r(R, 2 * I) for _ in range(k)], [torch.empty(R, I, device="cuda") for _ in range(k)], [r(R, I) for _ in range(k)], [torch.empty(R, 2 * I, device="cuda") for _ in range(k)]
    us = timed([ops.act_fwd(hs[i], ys[i], R, I, OPS.ACT_GEGLU) for i in range(k)])
    record(f"geglu_fwd {R}x{I}", us, 12 * R * I, 12 * R * I)
    us = timed([ops.act_bwd(ds[i], hs[i], dh[i], R, I, OPS.ACT_GEGLU) for i in range(k)])
    record(f"geglu_bwd {R}x{I}", us, 20 * R * I, 20 * R * I)
    del hs, ys, ds, dh
    n_el = 196608 * 512
    k = copies(n_el * 4)
    xs, ys, ds = [r(n_el) for _ in range(k)], [torch.empty(n_el, device="cuda") for _ in range(k)], [r(n_el) for _ in range(k)]
    us = timed([ops.act_fwd(xs[i], ys[i], 1, n_el, OPS.ACT_GELU) for i in range(k)])
    record("gelu_fwd 196608x512", us, 8 * n_el, 8 * n_el)
    us = timed([ops.act_bwd(ds[i], xs[i], ys[i], 1, n_el, OPS.ACT_GELU) for i in range(k)])
    record("gelu_bwd 196608x512", us, 12 * n_el, 12 * n_el)
    del xs, ys, ds
    # GridAttn's bilinear gather: V = 8 maps of 32 x 32 x 256, P = N HW D = 24576 points per view (the maps stay L2-resident; the rows stream)
    V, S, C, Pn = 8, 32, 256, 24576
    fm, xy = r(V, S, S, C), torch.rand(V, Pn, 2, device="cuda") * 2.2 - 1.1
    k = copies(V * Pn * C * 4)
    outs, dms = [torch.empty(V, Pn, C, device="cuda") for _ in range(k)], [torch.empty(V, S, S, C, device="cuda") for _ in range(k)]
    us = timed([ops.bilinear_gather_fwd(fm, xy, outs[i], V, S, S, C, Pn) for i in range(k)])
    record(f"bilinear_gather_fwd V{V} {S}x{S}x{C} P{Pn}", us, 4 * V * Pn * C + 4 * V * S * S * C, 4 * V * Pn * C + 4 * V * S * S * C)
    us = timed([ops.bilinear_gather_bwd(outs[i], xy, dms[i], V, S, S, C, Pn) for i in range(k)])
    record(f"bilinear_gather_bwd V{V} {S}x{S}x{C} P{Pn}", us, 4 * V * Pn * C + 4 * V * S * S * C, 4 * V * Pn * C + 8 * V * S * S * C)
    print(json.dumps({"what": "csrc/train.cu kernels, CUDA-event timed, buffers rotating through > 300 MB (HBM-resident inputs)", "hbm_peak_gbs": peak,
                      "peak_source": "MEASURED_PEAKS.json (copy bandwidth)", "gpu": torch.cuda.get_device_name(0), "rows": rows}))


if __name__ == "__main__":
    main()
