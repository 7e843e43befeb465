"""Per-process runtime glue: one NativeOps handle per CUDA device, weight-pack caches keyed on parameter versions.

There is no CPU execution path in the product: asking for ops on a non-CUDA device raises.  (The modules call
`runtime.get_ops(...)` through the module attribute, so tests/ can monkeypatch that one function with an emulating double of
the C library to check the host-side program logic on a GPU-less box; the package itself has no dispatch seam.)
"""
import torch

from .engine import PackedWeights
from .ops import MvdError, NativeOps

_OPS = {}


def get_ops(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise MvdError(
            f"mvdfusion_b200 runs on CUDA (sm_100a) only; got tensors on '{device}'. There is no CPU fallback — "
            "move the module and its inputs to a B200."
        )
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _OPS:
        _OPS[key] = NativeOps(torch.device("cuda", key))
    return _OPS[key]


def current_stream(device):
    if torch.device(device).type == "cuda":
        return torch.cuda.current_stream(device).cuda_stream
    return None


def params_signature(module):
    """Changes whenever a parameter / buffer is reassigned, moved or modified in place (load_state_dict, .cuda(), optimizer step)."""
    sig = []
    for t in list(module.parameters()) + list(module.buffers()):
        sig.append((t.data_ptr(), t._version))
    return hash(tuple(sig))


class PlanCache(dict):
    """A handful of compiled plans, least recently used first out.  A plan owns its arena, operand buffers, split-K workspace and a
    captured CUDA graph: calling sample() / apply_model() with ever-changing view counts or flags must not grow GPU memory without
    bound.  MVD_PLAN_CACHE in the environment sets the size (default 6)."""

    def __init__(self, capacity=None):
        super().__init__()
        import os
        self.capacity = capacity or int(os.environ.get("MVD_PLAN_CACHE", "6"))

    def __getitem__(self, key):
        value = super().pop(key)       # re-insert: dicts keep insertion order, so the first key is the least recently used one
        super().__setitem__(key, value)
        return value

    def __setitem__(self, key, value):
        if key in self:
            super().pop(key)
        super().__setitem__(key, value)
        while len(self) > self.capacity:
            super().pop(next(iter(self)))


class WeightCache:
    """PackedWeights of a module, rebuilt when the module's parameters change (packed weights are a derived cache)."""

    def __init__(self):
        self.sig = None
        self.state = None

    def get(self, module, ops):
        sig = (params_signature(module), id(ops))
        if sig != self.sig:
            self.state = {k: v.detach() for k, v in module.state_dict().items()}
            self.sig = sig
            self.packs = {}
            self.plans = PlanCache()
        return self.state

    def pack(self, module, ops, prefix=""):
        sd = self.get(module, ops)
        if prefix not in self.packs:
            self.packs[prefix] = PackedWeights(sd, ops, prefix)
        return self.packs[prefix]
