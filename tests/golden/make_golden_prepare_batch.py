"""Pins ViewFusion.prepare_batch to the reference's own implementation.

Runs only where /root/reference exists:   python tests/golden/make_golden_prepare_batch.py
It imports the unmodified reference (mvdfusion/viewfusion_zero_depth_rgb.py:165-273, utils/camera_utils.py:14-115) through
oracle/ref_shims, patches ONLY the two frozen encoders of the instance (`encode`, `encode_clip` -> the deterministic stand-ins of
tests/common.py; the VAE / CLIP weights are not available offline and are outside this row), calls the reference's
prepare_batch on a synthetic dataset batch and stores what it returns in tests/golden/prepare_batch_outputs.pt.
tests/test_boundary.py feeds the product's prepare_batch the same batch and compares.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MVD_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shims"), REF, ROOT, os.path.join(ROOT, "tests")]

from common import model_config, standin_clip_encode, standin_vae_encode, synthetic_dataset_batch  # noqa: E402

import mvdfusion.viewfusion_zero_depth_rgb as ref_vf  # noqa: E402  (the reference)


@torch.no_grad()
def main():
    cfg = model_config(64, 8, 1, 32)["params"]
    cfg["vae_config"] = {"target": "torch.nn.Identity"}
    cfg.pop("ddim_num_steps"), cfg.pop("latent_size")
    ref_vf.ViewFusion._init_clip = lambda self, clip_path: None
    ref = ref_vf.ViewFusion(**cfg).eval()
    ref.encode = lambda x: standin_vae_encode(x, ref.z_scale_factor)
    ref.encode_clip = lambda x: standin_clip_encode(x)
    out = {}
    for name, n_views, tc, seed in (("linspace", 9, {"input_batch_size": 1, "train_batch_size": 4, "random_views": False}, None),
                                    ("random", 7, {"input_batch_size": 1, "train_batch_size": 5, "random_views": True}, 123)):
        batch = synthetic_dataset_batch(n_views, 64, seed=3)
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        bl, bc, il, ic, cv = ref.prepare_batch(batch, tc, generator=gen)
        out[name] = {"n_views": n_views, "trainer_config": tc, "generator_seed": seed, "batch_latents": bl.clone(), "input_latents": il.clone(),
                     "clip_v_embed": cv.clone(), "batch_R": bc.R.clone(), "batch_T": bc.T.clone(), "batch_f": bc.focal_length.clone(),
                     "batch_p": bc.principal_point.clone(), "input_R": ic.R.clone(), "input_T": ic.T.clone()}
        print(name, tuple(bl.shape), tuple(il.shape), tuple(cv.shape), "input R == I:", bool(torch.allclose(ic.R[0], torch.eye(3), atol=1e-6)))
    path = os.path.join(HERE, "prepare_batch_outputs.pt")
    torch.save(out, path)
    print("wrote", path, f"({os.path.getsize(path) / 1e3:.1f} kB)")


if __name__ == "__main__":
    main()
