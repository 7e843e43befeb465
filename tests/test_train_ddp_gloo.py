"""CPU, world_size 2 over gloo: the training step under DistributedDataParallel as train.py:38,90-95 runs it (one scene per rank,
find_unused_parameters=True, gradients all-reduced — here with the bf16 compression hook bench.py --mode train registers).  Kernels
are emulated (tests/ops_double.py); the autograd Functions, the DDP hook-up of the facade's parameters and the collective are the
product's.  After backward both ranks hold the same gradients, equal to the mean of the two scenes' single-process gradients."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
TRAINER = {"input_batch_size": 1, "train_batch_size": 2, "random_views": False}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model_and_batch(scene_seed):
    from common import build_model, standin_clip_encode, standin_vae_encode, synthetic_dataset_batch
    import mvdfusion_b200.runtime as rt
    from ops_double import TorchOpsDouble
    dbl = TorchOpsDouble()
    if not hasattr(rt, "_real_get_ops"):
        rt._real_get_ops = rt.get_ops
    rt.get_ops = lambda dev: dbl
    m = build_model(64, 8, D=1, S=32)
    for p in m.parameters():
        p.requires_grad_(True)
    for p in m.view_attn.t_embedder.parameters():   # never read by forward
        p.requires_grad_(False)
    batch = synthetic_dataset_batch(3, 256, seed=scene_seed)
    images = batch.pop("images")
    batch["latents"] = standin_vae_encode(images, m.z_scale_factor) * 4.0
    batch["clip_embed"] = standin_clip_encode(images)
    m.train()
    return m, batch


def _loss(net, batch):
    torch.manual_seed(3)   # ViewFusion.forward draws t, the q_sample noise and GridAttn's depth jitter: the same draws on every rank and in the single-process runs
    return net(batch, TRAINER)


def _grads(m):
    return {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [os.path.dirname(HERE), HERE]
    torch.set_num_threads(2)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
    m, batch = _model_and_batch(20 + rank)
    net = torch.nn.parallel.DistributedDataParallel(m, find_unused_parameters=True)
    net.register_comm_hook(state=None, hook=default_hooks.bf16_compress_hook)
    loss = _loss(net, batch)
    loss.backward()
    torch.save({"loss": float(loss.detach()), "grads": _grads(m)}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_training_step_under_ddp_world2_averages_the_scene_gradients(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(str(tmp_path / f"rank{r}.pt")) for r in range(world)]
    assert got[0]["grads"].keys() == got[1]["grads"].keys() and len(got[0]["grads"]) > 100
    for k, g in got[0]["grads"].items():
        assert torch.equal(g, got[1]["grads"][k]), k          # one all-reduced gradient on every rank
    single = []
    for r in range(world):
        m, batch = _model_and_batch(20 + r)
        loss = _loss(m, batch)
        loss.backward()
        assert abs(float(loss.detach()) - got[r]["loss"]) <= 1e-4 * abs(float(loss.detach()))   # thread counts differ: summation order
        single.append(_grads(m))
    import mvdfusion_b200.runtime as rt
    rt.get_ops = rt._real_get_ops
    num = den = 0.0
    for k, g in got[0]["grads"].items():
        want = 0.5 * (single[0][k] + single[1][k])
        num += float((g - want).pow(2).sum())
        den += float(want.pow(2).sum())
    rel = (num / den) ** 0.5
    assert rel < 1e-2, rel   # bf16 compression of the all-reduce: 2^-9 relative per element
