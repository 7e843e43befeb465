"""CPU, world_size 2 over gloo: the view-sharded path (each rank denoises its own views, one all-gather of the 5-channel
latents per step) gives the same x_0 as the single-process run.  Kernels are emulated (tests/ops_double.py); the
partitioning, per-rank programs, noise slicing and the collective are the product's."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(n_views, steps, sharded, x_T="scene", stepwise=False, seed=None):
    """x_T: "scene" = the seeded scene's x_T for every rank; None = sample() draws it (each rank from its own RNG state);
    or an explicit tensor.  stepwise: the reference's side mode feed_prev_depth (DDIMSampler._sample_stepwise), made
    deterministic with eta = 0 and a zero DDIM noise so that the sharded and single-process RNG streams stay aligned."""
    from common import build_model, synthetic
    import mvdfusion_b200.runtime as rt
    from mvdfusion_b200.mvdfusion.cameras import PerspectiveCameras
    from ops_double import TorchOpsDouble
    dbl = TorchOpsDouble()
    if not hasattr(rt, "_real_get_ops"):
        rt._real_get_ops = rt.get_ops
    rt.get_ops = lambda dev: dbl  # test-side monkeypatch of the one function the modules call
    m = build_model(64, 8, D=1, S=32)
    m.ddim._make_schedule(steps, "uniform", 0.0 if stepwise else 1.0)
    if sharded:
        m.shard_views()
    sc = synthetic.scene_inputs(n_views, 32)
    de, dn = synthetic.step_noises(n_views, 1, 32, steps)
    cam = lambda c: PerspectiveCameras(c["R"], c["T"], c["f"], c["p"])
    xt = sc["x_T"] if isinstance(x_T, str) else x_T
    if seed is not None:
        torch.manual_seed(seed)
    if stepwise:
        m.ddim.feed_prev_depth = m.feed_prev_depth = True
        keep = torch.randn_like
        torch.randn_like = lambda t, **kw: torch.zeros_like(t)
        try:
            return m.ddim.sample(cam(sc["cams"]), sc["input_latents"], cam(sc["in_cams"]), sc["clip_v_embed"], unconditional_scale=2.5,
                                 depth=True, verbose=False, x_T=xt)
        finally:
            torch.randn_like = keep
    return m.ddim.sample(cam(sc["cams"]), sc["input_latents"], cam(sc["in_cams"]), sc["clip_v_embed"], unconditional_scale=2.5,
                         depth=True, verbose=False, x_T=xt, depth_eps=de, ddim_noise=dn)


def _worker(rank, world, port, n_views, steps, out_path, mode="plain"):
    sys.path[:0] = [os.path.dirname(HERE), HERE]
    torch.set_num_threads(2)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if mode == "own_xT":      # no x_T passed: every rank draws its own (different seeds) — the owners' rows must win
        x = _run(n_views, steps, sharded=True, x_T=None, seed=100 + rank)
    elif mode == "stepwise":  # the stepwise side mode under sharding (apply_model returns local views only)
        x = _run(n_views, steps, sharded=True, stepwise=True, seed=7)
    else:
        x = _run(n_views, steps, sharded=True)
    gathered = [torch.zeros_like(x) for _ in range(world)]
    dist.all_gather(gathered, x)
    for g in gathered[1:]:
        assert torch.equal(g, gathered[0])  # every rank ends with the same full set of views
    if rank == 0:
        torch.save(x, out_path)
    dist.destroy_process_group()


def test_view_sharding_world2_matches_single_process(tmp_path):
    n_views, steps = 4, 2
    out = str(tmp_path / "sharded.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_views, steps, out), nprocs=2, join=True)
    sharded = torch.load(out)
    single = _run(n_views, steps, sharded=False)
    import mvdfusion_b200.runtime as rt
    rt.get_ops = rt._real_get_ops
    rel = float((sharded - single).norm() / single.norm())
    assert rel < 2e-3, rel  # fp16-rounding noise of the emulated kernels (different batch shapes round differently); a mis-sliced view or noise row would be O(1)


def test_sharded_sample_without_xT_uses_the_owning_ranks_rows(tmp_path):
    """with views sharded and no x_T passed, each rank draws its own x_T; the rows of the rank that owns a view are
    the ones every rank must see from step 0 on — the result equals the single-process run started from that assembled x_T."""
    n_views, steps, world = 4, 2, 2
    out = str(tmp_path / "own.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_views, steps, out, "own_xT"), nprocs=world, join=True)
    sharded = torch.load(out)
    q = n_views // world
    rows = []
    for r in range(world):  # what rank r drew (sample() draws x_T first: torch.randn([B, 5, S, S]) on the model's device)
        torch.manual_seed(100 + r)
        rows.append(torch.randn([n_views, 5, 32, 32])[r * q:(r + 1) * q])
    single = _run(n_views, steps, sharded=False, x_T=torch.cat(rows))
    import mvdfusion_b200.runtime as rt
    rt.get_ops = rt._real_get_ops
    rel = float((sharded - single).norm() / single.norm())
    assert rel < 2e-3, rel


def test_sharded_stepwise_side_mode_matches_single_process(tmp_path):
    """DDIMSampler._sample_stepwise (feed_prev_depth / overwrite_x_noisy / condition drop) under view sharding: the
    local rows are updated and all-gathered every step instead of mixing (q, ...) and (B, ...) tensors."""
    n_views, steps, world = 4, 2, 2
    out = str(tmp_path / "stepwise.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_views, steps, out, "stepwise"), nprocs=world, join=True)
    sharded = torch.load(out)
    single = _run(n_views, steps, sharded=False, stepwise=True, seed=7)
    import mvdfusion_b200.runtime as rt
    rt.get_ops = rt._real_get_ops
    rel = float((sharded - single).norm() / single.norm())
    assert sharded.shape == single.shape and rel < 2e-3, rel
