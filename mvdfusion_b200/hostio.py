"""Host-side wire formats either side of the hot path: checkpoints, the on-disk outputs of demo.py, the GSO
evaluation data set with its fixed camera rig, and the resume-mid-epoch distributed sampler.  No kernels here: these are the
formats a user of the reference's demo.py / train.py finds unchanged.

  save_model / load_checkpoint   train.py:133-181 of the reference (dict keys, `{exp_dir}{ckpt_dir}/{mod}.pt`, strict=False resume)
  write_scene_outputs            demo.py:100-147 (…_eval_XXX_nN.jpg / .gif / _depth.png / _depth.npy / _depth.gif)
  GSO, look_at_rig               dataset/gso_test.py:19-160 (16 azimuths at 30 deg elevation, distance 1.5, focal 2.1875 NDC)
  StatefulDistributedSampler     utils/data_sampler_utils.py:10-143
  split_list, dict_to_device     utils/common_utils.py:72-83 and the batch mover used by demo.py:80-81
Images go through PIL (imageio / scikit-image, which the reference uses, are not part of this image).
"""
import glob
import json
import math
import os

import numpy as np
import torch
from torch.utils.data import Dataset, DistributedSampler


# ------------------------------------------------------------------------------------------------ small helpers
def split_list(a, n):
    """n nearly equal consecutive parts (utils/common_utils.py:72-83): how demo.py:63-64 spreads scenes over GPUs."""
    k, m = divmod(len(a), n)
    return [a[i * k + min(i, m):(i + 1) * k + min(i + 1, m)] for i in range(n)]


def dict_to_device(batch, to_device):
    return {k: (v.to(to_device) if torch.is_tensor(v) else v) for k, v in batch.items()}


def unnormalize(x):
    return torch.clip((x + 1.0) / 2.0, 0.0, 1.0)


def _bare(model):
    return model.module if hasattr(model, "module") else model  # DistributedDataParallel wrapper or the module itself


# ------------------------------------------------------------------------------------------------ checkpoints
def checkpoint_path(config, mod="latest"):
    return os.path.join(config["saver"]["exp_dir"] + config["saver"]["ckpt_dir"], f"{mod}.pt")


def save_model(config, model, optimizer, global_step, local_step, epoch, mod="latest"):
    """train.py:166-181: one file with the five keys the reference's load_model reads back."""
    path = checkpoint_path(config, mod)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({"local_step": local_step, "global_step": global_step, "epoch": epoch,
                "model_state_dict": _bare(model).state_dict(), "optimizer_state_dict": optimizer.state_dict()}, path)
    return path


def load_checkpoint(config, model, optimizer=None, mod="latest", map_location="cpu"):
    """train.py:141-161: resume when `{save_dir}/latest.pt` exists (strict=False, as the reference), else start from scratch.
    Returns (global_step, local_step, epoch)."""
    path = checkpoint_path(config, mod)
    if not os.path.exists(path):
        return 0, 0, 0
    ckpt = torch.load(path, map_location=map_location)
    _bare(model).load_state_dict(ckpt["model_state_dict"], strict=False)
    if optimizer is not None and "optimizer_state_dict" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return ckpt["global_step"], ckpt["local_step"], ckpt["epoch"]


# ------------------------------------------------------------------------------------------------ demo.py outputs
def _u8(a):
    return (np.asarray(a) * 255).astype(np.uint8)


def _save_gif(path, frames, duration_s=0.2):
    from PIL import Image
    ims = [Image.fromarray(f) for f in frames]
    ims[0].save(path, save_all=True, append_images=ims[1:], duration=int(duration_s * 1000), loop=0)


def write_scene_outputs(save_dir, global_step, val_idx, pred_rgb, gt_rgb, pred_latents, input_latents, batch_latents=None):
    """demo.py:100-147.  pred_rgb / gt_rgb: (B,3,H,W) in [0,1] (VAE-decoded); *_latents: (B,5,h,w) with the depth in channel 4.
    Writes  {step:07d}_eval_{idx:03d}_n{B}.jpg (predictions side by side), .gif (gt | prediction per frame), _depth.png and
    _depth.npy (input depth | predicted depths, [0,1]), _depth.gif.  Returns the paths."""
    from PIL import Image
    os.makedirs(save_dir, exist_ok=True)
    to_hwc = lambda t: t.detach().float().cpu().permute(0, 2, 3, 1).numpy()
    pred, gt = to_hwc(pred_rgb), to_hwc(gt_rgb)
    n = pred.shape[0]
    stem = os.path.join(save_dir, f"{int(global_step):07d}_eval_{int(val_idx):03d}_n{n}")
    paths = {"jpg": stem + ".jpg", "gif": stem + ".gif", "depth_png": stem + "_depth.png", "depth_npy": stem + "_depth.npy",
             "depth_gif": stem + "_depth.gif"}
    Image.fromarray(_u8(np.hstack(list(pred)))).save(paths["jpg"], quality=95)
    _save_gif(paths["gif"], [_u8(np.hstack((gt[j], pred[j]))) for j in range(n)])
    depth3 = lambda lat: to_hwc(unnormalize(lat[:, 4:]).clip(0.0, 1.0).expand(-1, 3, -1, -1))
    pred_d, in_d = depth3(pred_latents), depth3(input_latents)
    vis_depth = np.hstack((np.hstack(list(in_d)), np.hstack(list(pred_d))))
    Image.fromarray(_u8(vis_depth)).save(paths["depth_png"])
    with open(paths["depth_npy"], "wb") as fp:
        np.save(fp, vis_depth)
    _save_gif(paths["depth_gif"], [_u8(pred_d[j]) for j in range(n)])
    return paths


# ------------------------------------------------------------------------------------------------ GSO data set + camera rig
def look_at_rig(azimuth_rad, elevation_rad, distance=1.5):
    """World-to-view (R, T) of cameras on a sphere looking at the origin with up = +y, in pytorch3d's convention
    X_view = X_world @ R + T — what look_at_view_transform(dist, elev, azim (degrees), up=((0,1,0),)) returns
    (dataset/gso_test.py:134-141 calls it with azim = azimuths * 180/pi + 90).  azimuth_rad here is that final azimuth."""
    az, el = azimuth_rad.double(), elevation_rad.double()
    C = distance * torch.stack([torch.cos(el) * torch.sin(az), torch.sin(el), torch.cos(el) * torch.cos(az)], -1)
    z = torch.nn.functional.normalize(-C, dim=-1)
    up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64).expand_as(C)
    x = torch.nn.functional.normalize(torch.cross(up, z, dim=-1), dim=-1)
    y = torch.nn.functional.normalize(torch.cross(z, x, dim=-1), dim=-1)
    R = torch.stack([x, y, z], dim=-1)
    T = -torch.einsum("bji,bj->bi", R, C)
    return R.float(), T.float()


class GSO(Dataset):
    """dataset/gso_test.py:19-160: `{root}/{subset}.json` lists the scene folders; a scene holds 000.png … (RGBA renders on the fixed
    16-view rig).  __getitem__ returns the batch dict of README.md:87-96: images (16,3,S,S) in [0,1] on white, R, T, f, c,
    azimuth, elevation."""

    N_VIEWS = 16

    def __init__(self, root="", camera_type="fixed_set", stage="train", image_size=256, sample_batch_size=None, fix_elevation=True,
                 load_depth=False, load_mask=False, up_vec="y", subset="test"):
        super().__init__()
        if up_vec != "y":
            raise NotImplementedError("the shipped configs use up_vec = 'y' (configs/mvd_gso.yaml)")
        self.root, self.camera_type, self.stage, self.image_size = root, camera_type, stage, image_size
        listing = f"{root}/{subset}.json"
        assert os.path.exists(listing), "subset not found"
        with open(listing) as fp:
            self.subset_list = json.load(fp)
        self.azimuths = torch.arange(self.N_VIEWS, dtype=torch.float32) * (2 * math.pi / self.N_VIEWS)
        self.elevations = torch.full((self.N_VIEWS,), math.pi / 6)
        self.R, self.T = look_at_rig(self.azimuths + math.pi / 2, self.elevations, 1.5)
        self.f = torch.full((self.N_VIEWS, 2), 2.1875)
        self.c = torch.zeros(self.N_VIEWS, 2)

    def __len__(self):
        return len(self.subset_list)

    def _load_images(self, scene_dir, idxs):
        from PIL import Image
        out = []
        for i in idxs:
            im = Image.open(f"{scene_dir}/{int(i):03d}.png").convert("RGBA").resize((self.image_size, self.image_size), Image.BILINEAR)
            a = torch.from_numpy(np.asarray(im, dtype=np.float32) / 255.0)
            rgb, alpha = a[..., :3].clone(), a[..., 3:]
            rgb[(alpha < 0.5).expand_as(rgb)] = 1.0   # white background (gso_test.py:100-104)
            out.append(rgb)
        return torch.stack(out).permute(0, 3, 1, 2).contiguous()

    def __getitem__(self, index):
        scene_dir = f"{self.root}/{self.subset_list[index]}/"
        if self.camera_type == "fixed_set":
            assert len(glob.glob(scene_dir + "*.png")) == 32
        idx = torch.arange(0, self.N_VIEWS)
        return {"index": index, "idx": self.subset_list[index], "images": self._load_images(scene_dir, idx), "R": self.R[idx],
                "T": self.T[idx], "f": self.f[idx], "c": self.c[idx], "azimuth": self.azimuths[idx], "elevation": self.elevations[idx]}


# ------------------------------------------------------------------------------------------------ sampler
class StatefulDistributedSampler(DistributedSampler):
    """utils/data_sampler_utils.py:10-143: torch's DistributedSampler that can resume in the middle of an epoch — the first
    `start_iter * batch_size` indices of this rank are skipped (train.py:43-47 passes start_iter = local_step of the checkpoint);
    set_epoch(epoch, zero_start=True) clears the offset for the following epochs."""

    def __init__(self, dataset, num_replicas=None, rank=None, shuffle=True, seed=0, drop_last=False, start_iter=0, batch_size=1):
        super().__init__(dataset, num_replicas=num_replicas, rank=rank, shuffle=shuffle, seed=seed, drop_last=drop_last)
        self.start_iter, self.batch_size = start_iter, batch_size

    def __iter__(self):
        return iter(list(super().__iter__())[self.start_iter * self.batch_size:])

    def set_epoch(self, epoch, zero_start=True):
        super().set_epoch(epoch)
        if zero_start:
            self.start_iter = 0
