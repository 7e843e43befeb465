"""SURVEY.md §8d parity gate, second half: "state separately the figure vs the reference under bf16 / fp16 autocast".

How far does the REFERENCE ALGORITHM itself move when its contractions run at reduced precision the stock way
(torch.autocast around the fp32 oracle: bf16 / fp16 operands, fp32 accumulate inside the library kernels, fp32 norms / softmax by the
autocast policy)?  That is the context for the product's own figure (fp16 operands, fp32 accumulate, fp32 residual stream, hi/lo split
operands on the stem / head / skips: 6.7e-4 .. 7.1e-4 rel-L2 at full size, gate 1e-3).

Runs on the host cores (the oracle is CPU code; the comparison is between arithmetic types, not devices); full-size UNet (1.03 B
parameters), N = 2 views, cfg 2.5, the same seeds / timesteps as tests/test_gpu_parity.py::test_apply_model_full_size_vs_oracle.
Test infrastructure: imports oracle/, never imported by the product.

    python tests/autocast_parity.py --out profiles/r02_autocast_parity.json
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--channels", type=int, default=320)
    ap.add_argument("--views", type=int, default=2)
    ap.add_argument("--timesteps", type=int, nargs="+", default=[781, 21])
    a = ap.parse_args()
    from common import build_model, rel_l2, state_dict_cpu, unet_cfg_of
    from mvdfusion_b200 import synthetic
    from oracle import mvd_oracle as O

    N, S, D = a.views, 32, 1
    m = build_model(a.channels, 8, D=D, S=S)
    sd = state_dict_cpu(m)
    sc = synthetic.scene_inputs(N, S, seed=3)
    de, _ = synthetic.step_noises(N, D, S, 1, seed=4)
    rows = []
    for tv in a.timesteps:
        t = torch.full((N,), tv, dtype=torch.long)
        run = lambda: O.apply_model(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], t, de[0],
                                    unet_cfg=unet_cfg_of(m), D=D, cfg_scale=2.5)
        t0 = time.time()
        with torch.no_grad():
            ref = run()
        rec = {"t": tv, "fp32_seconds": round(time.time() - t0, 1)}
        for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
            try:
                # torch.cross has no mixed-type CPU kernel (camera centre in fp32 x autocast direction): promote its operands
                cross = torch.cross
                torch.cross = lambda u, v, dim=-1: cross(u.float(), v.float(), dim=dim)
                try:
                    with torch.no_grad(), torch.autocast("cpu", dtype=dt):
                        y = run()
                finally:
                    torch.cross = cross
                rec[f"autocast_{name}_rel_l2_vs_fp32"] = rel_l2(y.float(), ref)
            except Exception as e:  # an op without a CPU kernel for the type
                rec[f"autocast_{name}_rel_l2_vs_fp32"] = None
                rec[f"autocast_{name}_error"] = str(e).splitlines()[0][:200]
        print(rec, flush=True)
        rows.append(rec)
    out = {"what": "fp32 oracle (reference algorithm) under torch.autocast on the host cores vs the same oracle in fp32; full-size UNet, "
                   f"N = {N} views, cfg 2.5, scene seed 3 / noise seed 4", "model_channels": a.channels, "rows": rows,
           "product_figure": "6.7e-4 .. 7.1e-4 rel-L2 vs the fp32 oracle at full size (profiles/r02_parity_v6.jsonl), gate 1e-3"}
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
