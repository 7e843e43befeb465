#!/usr/bin/env python
"""Phase timeline of the fused GridAttn transformer kernel (csrc/dit.cu), first tile of every CTA; needs the instrumented library
(`make -C mvdfusion_b200/csrc trace`, MVD_B200_LIB=mvdfusion_b200/libmvd_b200_trace.so).  B200 only."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    from mvdfusion_b200 import _lib, ops as OPS
    from test_gpu_ops import _dit_inputs
    lib = _lib.load()
    nat = OPS.NativeOps("cuda:0")
    R, V, nl = 65536, 8, 3
    t, layers = _dit_inputs(R, V, nl)
    g = {k: v.cuda() for k, v in t.items()}
    gl = [{k: v.cuda() for k, v in l.items()} for l in layers]
    call = nat.gridattn_dit(g["tokens"], 736, g["w_pre"], g["b_pre"], gl, g["pool_w"], g["pool_b"], g["pooled"], R, V, 1e-6)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        call(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(st)
    e1.record()
    torch.cuda.synchronize()
    print(f"kernel: {e0.elapsed_time(e1) * 1e3:.1f} us for {R} rows ({R // 128} tiles)")
    if not hasattr(lib, "mvd_debug_dit_trace"):
        return
    buf = np.zeros(160 * 64, dtype=np.uint32)
    lib.mvd_debug_dit_trace.argtypes = [ctypes.c_void_p]
    lib.mvd_debug_dit_trace(ctypes.c_void_p(buf.ctypes.data))
    tr = buf.reshape(160, 64)[:148].astype(np.int64)
    rel = ((tr - tr[:, :1]) & 0xFFFFFFFF) / 1965.0  # us
    med = np.median(rel, axis=0)
    names = {0: "start", 1: "x_ready", 2: "gelu done", 40: "last x_done", 41: "pool done"}
    for l in range(nl):
        names.update({3 + 8 * l: f"L{l} begin", 4 + 8 * l: f"L{l} x_done(fc2 prev)", 5 + 8 * l: f"L{l} LN1 done", 6 + 8 * l: f"L{l} attention done",
                      7 + 8 * l: f"L{l} x_done(proj)", 8 + 8 * l: f"L{l} LN2 done", 9 + 8 * l: f"L{l} MLP epilogues done"})
    prev = 0.0
    for i in sorted(names):
        print(f"{names[i]:28s} {med[i]:8.2f} us   (+{med[i] - prev:6.2f})")
        prev = med[i]


if __name__ == "__main__":
    main()
