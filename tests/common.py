"""Shared helpers of the test-suite (model configs, building the product model, rel-L2)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mvdfusion_b200 import synthetic  # noqa: E402


def unet_params(model_channels=320, num_heads=8, image_size=32):
    """configs/mvd_gso.yaml:30-46 of the reference (model_channels / heads reducible for small-size parity cases)."""
    return dict(image_size=image_size, in_channels=10, out_channels=5, model_channels=model_channels,
                attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=num_heads,
                use_spatial_transformer=True, use_view_aligned_transformer=True, transformer_depth=1, context_dim=768,
                use_checkpoint=True, legacy=False)


def model_config(model_channels=320, num_heads=8, D=1, S=32, drop_conditions=False, ddim_steps=50):
    """model: section of configs/mvd_gso.yaml (target strings exactly as the reference writes them)."""
    return {
        "target": "mvdfusion.viewfusion_zero_depth_rgb.ViewFusion",
        "params": {
            "vae_path": None, "clip_path": None, "unet_path": None, "z_scale_factor": 0.18215, "objective": "noise",
            "loss_type": "l2", "embed_camera_pose": True, "finetune_projection": True, "finetune_unet": False,
            "finetune_cross_attn": True, "finteune_view_attn": True, "drop_conditions": drop_conditions,
            "ddim_num_steps": ddim_steps, "latent_size": S,
            "view_attn_config": {"target": "mvdfusion.view_attn_efficient2.GridAttn",
                                 "params": {"in_channels": 5, "input_size": S, "output_dim": 768, "num_layers": 3,
                                            "z_near_far_scale": 0.8, "n_pts_per_ray": D}},
            "unet_config": {"target": "mvdfusion.unet.UNetModel", "params": unet_params(model_channels, num_heads, S)},
            "ddpm_config": {"target": "mvdfusion.scheduler.DDPMScheduler", "params": {"timesteps": 1000}},
            "vae_config": None,
        },
    }


def build_model(model_channels=320, num_heads=8, D=1, S=32, seed=1234, device="cpu", **kw):
    from mvdfusion_b200.config import instantiate_from_config
    m = instantiate_from_config(model_config(model_channels, num_heads, D, S, **kw))
    synthetic.randomize_parameters(m, seed)
    return m.to(device).eval()


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def state_dict_cpu(model):
    return {k: v.detach().float().cpu() if v.is_floating_point() else v.detach().cpu() for k, v in model.state_dict().items()}


def unet_cfg_of(model):
    um = model.unet_model.unet_model
    return {"model_channels": um.model_channels, "num_heads": um.num_heads, "image_size": um.image_size,
            "channel_mult": list(um.channel_mult)}


def record_parity(name, value, tol):
    """Append a measured rel-L2 to gpurun_out/parity.jsonl (brought back from the GPU box) and echo it."""
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, "rel_l2": value, "tolerance": tol}) + "\n")
    except OSError:
        pass
    print(f"parity {name}: rel-L2 {value:.3e} (tolerance {tol:g})")
    return value


# ---- deterministic stand-ins for the frozen encoders (outside the hot path; weights not available offline), shared by
#      tests/golden/make_golden_prepare_batch.py (patched onto the REFERENCE's ViewFusion) and tests/test_boundary.py
def standin_vae_encode(images, z_scale=0.18215):
    """(S,3,H,W) in [0,1] -> (S,4,H/8,W/8): 8x8 area pooling of the normalised image + a 4th channel (their mean), x z_scale"""
    x = torch.clip(images * 2 - 1.0, -1.0, 1.0)
    p = torch.nn.functional.avg_pool2d(x, 8)
    return torch.cat([p, p.mean(dim=1, keepdim=True)], dim=1) * z_scale


def standin_clip_encode(images):
    """(S,3,H,W) -> (S,1,768)"""
    m = images.mean(dim=(1, 2, 3)).reshape(-1, 1, 1)
    return torch.sin(torch.linspace(0.0, 20.0, 768).reshape(1, 1, 768) * (1.0 + m))


def synthetic_dataset_batch(n_views=9, side=64, seed=0):
    """A dataset-style batch (README.md:87-96 of the reference): images, depths, R, T, f, c of `n_views` look-at cameras in an
    arbitrarily rotated world frame (so that the relative-camera step has something to undo)."""
    g = torch.Generator().manual_seed(seed)
    R, T, f, p = synthetic.gso_rig(n_views - 1)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return {"images": torch.rand(n_views, 3, side, side, generator=g), "depths": torch.rand(n_views, 1, side, side, generator=g),
            "R": torch.einsum("ij,bjk->bik", q, R).contiguous(), "T": T.clone(), "f": f.clone(), "c": p.clone()}


class fast_init:
    """Context manager: parameter initialisers become no-ops (shapes / names are what a structural test needs; drawing 1.1 B random
    numbers takes most of a minute on the CPU)."""
    NAMES = ("kaiming_uniform_", "kaiming_normal_", "uniform_", "normal_", "trunc_normal_", "xavier_uniform_", "xavier_normal_")

    def __enter__(self):
        import torch.nn.init as I
        self.saved = {n: getattr(I, n) for n in self.NAMES}
        for n in self.NAMES:
            setattr(I, n, lambda t, *a, **k: t)
        return self

    def __exit__(self, *exc):
        import torch.nn.init as I
        for n, f in self.saved.items():
            setattr(I, n, f)


def build_cpu_shim(sources, out_dir, name="libmvd_cpuemul.so"):
    """TEST INFRASTRUCTURE: compile kernel sources of mvdfusion_b200/csrc as plain C++ against tests/native/cpu_emul/cuda_on_cpu.h
    (-DMVD_CPU_EMULATION: one host thread per CUDA thread of a block) into one shared object; returns its path."""
    import subprocess
    out = os.path.join(str(out_dir), name)
    objs = []
    for src in sources:
        obj = os.path.join(str(out_dir), os.path.basename(src) + ".o")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-x", "c++", "-DMVD_CPU_EMULATION", "-I", os.path.join(ROOT, "tests", "native", "cpu_emul"),
                               "-fPIC", "-c", os.path.join(ROOT, "mvdfusion_b200", "csrc", src), "-o", obj])
        objs.append(obj)
    subprocess.check_call(["g++", "-shared", "-o", out] + objs + ["-lpthread"])
    return out


def guarded_clone(t):
    """TEST INFRASTRUCTURE (memcheck stand-in for kernels run on the CPU shim): a copy of `t` whose last byte sits right in front of an
    inaccessible page (anonymous mmap + mprotect(PROT_NONE)) — a kernel that reads or writes past the end of the buffer dies with SIGSEGV
    instead of silently touching a neighbour.  The start stays 16-byte aligned (sizes are rounded up to 16 bytes: <= 15 bytes of slack)."""
    import ctypes
    import mmap
    page = mmap.PAGESIZE
    t = t.detach().contiguous()
    nbytes = max(16, (t.numel() * t.element_size() + 15) // 16 * 16)
    size = (nbytes + page - 1) // page * page + page
    m = mmap.mmap(-1, size)
    probe = ctypes.c_char.from_buffer(m)
    base = ctypes.addressof(probe)
    del probe
    libc = ctypes.CDLL(None, use_errno=True)
    libc.mprotect.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    if libc.mprotect(base + size - page, page, 0) != 0:
        raise OSError(ctypes.get_errno(), "mprotect failed")
    off = size - page - nbytes
    g = torch.frombuffer(m, dtype=t.dtype, count=t.numel(), offset=off).reshape(t.shape)
    g.copy_(t)
    return g


def shim_ops(lib_path, monkeypatch):
    """the product's own binding layer (ops.NativeOps: argument order, workspace sizes, job tables) over a CPU-shim build of the kernels;
    ops._ptr is patched to accept host tensors for the duration of the test"""
    import ctypes
    from mvdfusion_b200 import _lib, ops as OPS
    lib = ctypes.CDLL(lib_path)
    for name, argtypes in _lib.SIGNATURES.items():
        if hasattr(lib, name):
            getattr(lib, name).argtypes = argtypes
            getattr(lib, name).restype = _lib._RESTYPE.get(name, ctypes.c_int32)

    class ShimOps(OPS.NativeOps):
        def __init__(self):
            self.lib, self.device = lib, torch.device("cpu")

        # every buffer the kernels see ends at a guard page (see guarded_clone)
        guarded = staticmethod(guarded_clone)

        def empty(self, shape, dtype):
            return guarded_clone(torch.full(shape, float("nan") if dtype.is_floating_point else 255, dtype=dtype))

        def zeros(self, shape, dtype):
            return guarded_clone(torch.zeros(shape, dtype=dtype))

    def ptr(t, dtype=None):
        if t is None:
            return None
        if dtype is not None and t.dtype != dtype:
            raise OPS.MvdError(f"expected {dtype}, got {t.dtype}")
        return t.data_ptr()

    monkeypatch.setattr(OPS, "_ptr", ptr)
    return ShimOps()
