"""CPU: the host-side wire formats of SURVEY.md §8f rank 4 (mvdfusion_b200/hostio.py) — checkpoints, demo.py's on-disk outputs, the
GSO data set + camera rig, the resume-mid-epoch sampler."""
import json
import math
import os

import numpy as np
import torch

from mvdfusion_b200 import hostio


def test_split_list_matches_demo_partitioning():
    parts = hostio.split_list(torch.arange(20), 8)
    assert [len(p) for p in parts] == [3, 3, 3, 3, 2, 2, 2, 2] and torch.equal(torch.cat(parts), torch.arange(20))


def test_checkpoint_round_trip_with_the_reference_keys(tmp_path):
    cfg = {"saver": {"exp_dir": str(tmp_path) + "/", "ckpt_dir": "checkpoints/"}}
    model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    model(torch.randn(2, 4)).sum().backward()
    opt.step()
    path = hostio.save_model(cfg, torch.nn.parallel.DataParallel(model), opt, global_step=120, local_step=20, epoch=3)  # .module wrapper as in train.py
    assert path == f"{tmp_path}/checkpoints/latest.pt"
    ck = torch.load(path)
    assert set(ck) == {"local_step", "global_step", "epoch", "model_state_dict", "optimizer_state_dict"}  # train.py:171-177
    m2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    o2 = torch.optim.AdamW(m2.parameters(), lr=1e-3)
    assert hostio.load_checkpoint(cfg, m2, o2) == (120, 20, 3)
    assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), m2.state_dict().values()))
    assert o2.state_dict()["state"][0]["step"] == opt.state_dict()["state"][0]["step"]
    assert hostio.load_checkpoint({"saver": {"exp_dir": str(tmp_path) + "/none/", "ckpt_dir": "c/"}}, m2) == (0, 0, 0)


def test_scene_outputs_follow_demo_py(tmp_path):
    from PIL import Image
    g = torch.Generator().manual_seed(0)
    B = 3
    pred, gt = torch.rand(B, 3, 16, 16, generator=g), torch.rand(B, 3, 16, 16, generator=g)
    lat = torch.rand(B, 5, 4, 4, generator=g) * 2 - 1
    inp = torch.cat([torch.rand(1, 4, 4, 4, generator=g), torch.zeros(1, 1, 4, 4)], 1)
    p = hostio.write_scene_outputs(str(tmp_path / "vis"), 1234, 7, pred, gt, lat, inp)
    assert os.path.basename(p["jpg"]) == "0001234_eval_007_n3.jpg"
    assert Image.open(p["jpg"]).size == (16 * B, 16)
    gif = Image.open(p["gif"])
    assert gif.n_frames == B and gif.size == (32, 16) and gif.info["duration"] == 200
    depth = np.load(p["depth_npy"])
    want = torch.cat([hostio.unnormalize(inp[:, 4:]), hostio.unnormalize(lat[:, 4:])], 0)  # input depth | predicted depths, side by side
    want = torch.cat(list(want.expand(-1, 3, -1, -1).permute(0, 2, 3, 1)), dim=1).numpy()
    assert depth.shape == (4, 4 * (B + 1), 3) and np.allclose(depth, want)
    assert np.array_equal(np.asarray(Image.open(p["depth_png"])), (want * 255).astype(np.uint8))  # png is lossless
    assert Image.open(p["depth_gif"]).n_frames == B


def test_gso_dataset_and_rig(tmp_path):
    from PIL import Image
    root = tmp_path / "gso"
    (root / "scene_a").mkdir(parents=True)
    json.dump(["scene_a"], open(root / "test.json", "w"))
    for i in range(32):
        a = np.zeros((32, 32, 4), np.uint8)
        a[8:24, 8:24] = (200, 100 + i, 50, 255)   # opaque square on a transparent background
        Image.fromarray(a, "RGBA").save(root / "scene_a" / f"{i:03d}.png")
    ds = hostio.GSO(root=str(root), image_size=16, subset="test")
    b = ds[0]
    assert len(ds) == 1 and b["images"].shape == (16, 3, 16, 16) and b["R"].shape == (16, 3, 3) and b["f"].shape == (16, 2)
    assert torch.allclose(b["images"][:, :, 0, 0], torch.ones(16, 3))                   # transparent -> white
    assert torch.allclose(b["images"][3, :, 8, 8], torch.tensor([200, 103, 50]) / 255.0, atol=1e-6)
    # the rig (dataset/gso_test.py:116-149): orthonormal, right-handed, centres at distance 1.5 and 30 deg elevation, looking at the origin
    R, T = b["R"], b["T"]
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(16, 3, 3), atol=1e-6) and torch.allclose(torch.det(R), torch.ones(16), atol=1e-6)
    C = -torch.einsum("bj,bij->bi", T, R)                                                # camera centres = -T R^T
    assert torch.allclose(C.norm(dim=1), torch.full((16,), 1.5), atol=1e-6) and torch.allclose(C[:, 1], torch.full((16,), 1.5 * math.sin(math.pi / 6)), atol=1e-6)
    origin_view = T                                                                      # 0 @ R + T
    assert torch.allclose(origin_view[:, :2], torch.zeros(16, 2), atol=1e-6) and torch.allclose(origin_view[:, 2], torch.full((16,), 1.5), atol=1e-6)
    assert float(b["f"][0, 0]) == 2.1875 and torch.allclose(b["azimuth"][1] - b["azimuth"][0], torch.tensor(math.pi / 8))


def test_stateful_sampler_resumes_mid_epoch():
    data = list(range(23))
    full = hostio.StatefulDistributedSampler(data, num_replicas=2, rank=1, shuffle=True, seed=5)
    full.set_epoch(2)
    ref = torch.utils.data.DistributedSampler(data, num_replicas=2, rank=1, shuffle=True, seed=5)
    ref.set_epoch(2)
    order = list(ref)
    assert list(full) == order and len(full) == len(order)
    resumed = hostio.StatefulDistributedSampler(data, num_replicas=2, rank=1, shuffle=True, seed=5, start_iter=4, batch_size=2)
    resumed.set_epoch(2, zero_start=False)       # resume inside epoch 2 after 4 iterations of 2 samples
    assert list(resumed) == order[8:]
    resumed.set_epoch(3)                          # the next epoch starts from the top again
    ref.set_epoch(3)
    assert list(resumed) == list(ref)
