"""mvdfusion_b200 — B200-native (sm_100a) implementation of MVD-Fusion's multi-view denoising hot path.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/mvd_b200.h) -> libmvd_b200.so
  _lib, ops        ctypes binding / tensor-level bound calls
  engine, denoise  weight packing, buffer arena, kernel-call programs, CUDA-graph step plan
  mvdfusion/       host-side mirror of the reference's Python surface (same class names, constructor
                   arguments and parameter names as zhizdev/mvdfusion's `mvdfusion` package)
  compat           registers the mirror under the reference's module paths (yaml `target:` strings)
"""
__version__ = "0.1.0"
