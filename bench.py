#!/usr/bin/env python
"""Benchmark of the multi-view denoising hot path (BASELINE.json metric: UNet denoising steps/sec, N-view, 256^2, DDIM).

    python bench.py [--gpus N --steps K --warmup W]          # this repo's sm_100a path, one JSON line
    python bench.py --impl reference [...]                   # the reference algorithm on the host CPUs (oracle port)

One *step* = one iteration of DDIMSampler.sample's loop for one scene of `--views` views at 256^2 (32x32x(4+1) latents):
GridAttn over all views + UNet over the views (x2: classifier-free guidance 2.5, both branches in one batch) + CFG
combine + DDIM update.  Workload at N=1 = BASELINE.json configs[1]: N=8 views, 1xB200, full-size UNet
(1.034 B parameters, random-init — no checkpoints offline), D = 1 depth sample per ray, synthetic GSO camera rig.

Multi-GPU (`--gpus G`, one process per GPU under torchrun), both of SURVEY.md §8(e)'s modes are measured in the same run:
  * replicas (the headline `value`; weak scaling): one independent scene of `--views` views per rank, no data-path
    collective — what the reference's demo.py does (demo.py:63-64), and the throughput mode;
  * view-sharded (reported under `"sharded"`; strong scaling, the latency mode): the views of ONE scene are split over the
    G ranks — every rank runs GridAttn for its own query views against all views and the UNet on its own views, then the
    ranks exchange the updated 5-channel latents with one NCCL all-gather per step.
`--mode shard` makes the sharded figure the headline instead.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

F_UNET_GFLOP = 225.1       # per view-pass, 32x32 latents, D=1
F_GRID_GFLOP = 3.590e-3    # x N^2 S^2 D
METRIC = "UNet denoising steps/sec (N-view, 256^2, 50-step DDIM)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--train-graph", action="store_true",
                    help="--mode train on one GPU: capture forward + backward + AdamW of the training step in ONE CUDA graph and time its replays "
                         "(the eager step is host-bound); experimental")
    ap.add_argument("--mode", default="replicas", choices=["shard", "replicas", "train"],
                    help="train: BASELINE configs[3] — forward + backward + AdamW of one scene per rank per step under DDP (not the headline metric)")
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--latent", type=int, default=32, choices=[32, 64],
                    help="latent side: 32 = 256^2 images (the headline workload), 64 = 512^2 (BASELINE configs[4], HBM-bandwidth stress)")
    ap.add_argument("--cfg", type=float, default=2.5)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-table", default="", help="write the per-kernel time table (json) here")
    ap.add_argument("--reps", type=int, default=0, help="repetitions of the timed K-step loop (0 = enough for >= 2 s, at least 3); the median is reported")
    return ap.parse_args()


def step_flops(n_views, S, D, cfg):
    p = 1 if cfg == 1.0 else 2
    f_unet = F_UNET_GFLOP if S == 32 else 1047.1  # F_U(64, 1)
    return (F_GRID_GFLOP * n_views * n_views * S * S * D + p * n_views * f_unet) * 1e9


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------ CPU arms
def workload_string(args):
    """`config.workload`, identical in both arms (the driver compares the strings)."""
    S = args.latent
    return (f"N={args.views} views {8 * S}^2 ({S}x{S}x5 latents) DDIM step, cfg {args.cfg} (cond+uncond), D=1, full-size UNet 1.034B params "
            f"random-init + GridAttn (BASELINE configs[{1 if S == 32 else 4}])")


def cpu_arm(args, steps, warmup, budget_s):
    """The reference's CPU implementation of the path on the host cores, all threads: `steps` timed + `warmup` untimed FULL
    denoising steps (all `views` views, both CFG passes, GridAttn, DDIM update) of the same workload as the native arm.
    kind "reference": the reference's own modules (sourceless bytecode in oracle/_ref, built by oracle/build_ref.py from
    /root/reference) through oracle/ref_shims; kind "port": oracle/mvd_oracle.py when oracle/_ref is absent.
    Returns (cpu_baseline dict, seconds per step, steps actually timed)."""
    from common import build_model, model_config, state_dict_cpu, synthetic, unet_cfg_of
    from oracle import ref_runner as RR
    torch.set_num_threads(os.cpu_count())
    n, S, D = args.views, args.latent, 1
    sc = synthetic.scene_inputs(n, S)
    if RR.available():
        kind = "reference"
        m = RR.build_reference_model(model_config(320, 8, D, S)["params"])
        torch.manual_seed(1)
        times, _ = RR.time_denoising_steps(m, sc, args.cfg, steps, warmup, budget_s=budget_s)
        how = "the reference's own ViewFusion / DDIMSampler.denoise_apply (oracle/_ref bytecode through oracle/ref_shims)"
    else:
        from oracle import mvd_oracle as O
        kind = "port"
        m = build_model(320, 8, D=D, S=S)
        sd, ucfg = state_dict_cpu(m), unet_cfg_of(m)
        de, dn = synthetic.step_noises(n, D, S, 1)
        tab = O.ddim_tables(sd["scheduler.alphas_cumprod"], 50, 1.0)
        times, t_start = [], time.perf_counter()
        with torch.no_grad():
            for i in range(warmup + steps):
                index = 49 - (i % 50)
                t0 = time.perf_counter()
                t = torch.full((n,), int(tab["timesteps"][index]), dtype=torch.long)
                eps = O.apply_model(sd, sc["x_T"], sc["cams"], sc["input_latents"], sc["in_cams"], sc["clip_v_embed"], t, de[0],
                                    unet_cfg=ucfg, D=D, cfg_scale=args.cfg)
                O.ddim_update(sc["x_T"], eps, tab, index, dn[0])
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
                    if time.perf_counter() - t_start > budget_s:
                        break
        how = "oracle/mvd_oracle.py (CPU port of the reference algorithm; oracle/_ref was not built)"
    step_s = sum(times) / len(times)
    return {"value": 1.0 / step_s, "unit": "steps/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{len(times)} full denoising steps (all {n} views, {'2 UNet passes' if args.cfg != 1.0 else '1 UNet pass'} + GridAttn + DDIM "
                      f"update) after {warmup} warm-up, fp32, {how}"}, step_s, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = float(os.environ.get("MVD_REF_BUDGET_S", "300"))
    cb, step_s, timed = cpu_arm(args, max(1, args.steps), max(0, args.warmup), budget)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "steps/s", "n_gpus": args.gpus, "steps": timed,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args), "steps_requested": args.steps,
                       "note": "the reference's CPU implementation on the host cores of this box (rank 0 only); `steps` = the steps "
                               f"actually timed (stops early once {budget:g} s of wall time are spent)"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 7]
        os.unlink(self.f.name)
        sm, mx, reasons = [], 0, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ per-kernel timing
def kernel_table(prog, stream):
    """Per-launch device time of every bound call of one step, CUDA events on the launching stream; a spin kernel is
    queued first so that the host runs ahead of the device and the events bracket back-to-back kernel execution."""
    calls = prog.calls
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in calls]
    torch.cuda.synchronize()
    torch.cuda._sleep(int(1.2e8))
    for c, (e0, e1) in zip(calls, ev):
        e0.record()
        c(stream)
        e1.record()
    torch.cuda.synchronize()
    agg = {}
    for c, (e0, e1) in zip(calls, ev):
        name = c.meta.get("kernel", c.name.replace("mvd_", ""))
        a = agg.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        a["calls"] += 1
        a["ms"] += e0.elapsed_time(e1)
        a["flops"] += c.meta.get("flops", 0.0)
        a["bytes"] += c.meta.get("bytes", 0.0)
    total = sum(a["ms"] for a in agg.values())
    rows = []
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        rows.append({"kernel": name, "calls": a["calls"], "ms": round(a["ms"], 4), "share": round(a["ms"] / total, 4),
                     "tflops": round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 2) if a["flops"] else None,
                     "algo_gb": round(a["bytes"] / 1e9, 4) if a["bytes"] else None})
    return rows, total


def kernel_graph_time(prog, kernel_name, stream_obj, reps=5):
    """Average device time of the launches of ONE kernel of the step, replayed back to back from a CUDA graph that holds only
    those launches (same arguments, same buffers, same order as in the step; programmatic dependent launch as in the step):
    the kernel's duration without the per-launch event overhead of kernel_table().  Returns (ms per step's worth, launches)."""
    calls = [c for c in prog.calls if c.meta.get("kernel", c.name.replace("mvd_", "")) == kernel_name]
    if not calls:
        return None, 0
    for c in calls:
        c(stream_obj.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for c in calls:
            c(torch.cuda.current_stream().cuda_stream)
    g.replay()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, len(calls)


# ------------------------------------------------------------------------------------------------ native arm
def run_native(args):
    import torch.distributed as dist
    from common import build_model, synthetic
    from mvdfusion_b200 import _lib
    from mvdfusion_b200.mvdfusion.cameras import PerspectiveCameras
    from mvdfusion_b200.runtime import current_stream

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, S, D, K, Wm = args.views, args.latent, 1, args.steps, args.warmup
    shard = world > 1 and args.mode == "shard"
    if world > 1 and n % world:
        raise SystemExit(f"--views {n} does not shard over {world} ranks")

    model = build_model(320, 8, D=D, S=S, device=dev)
    total = K + Wm
    de1, dn1 = synthetic.step_noises(n, D, S, min(total, 50), seed=1)
    idx = [i % de1.shape[0] for i in range(total)]
    de, dn = de1[idx], dn1[idx]
    rows = torch.stack([model.ddim.step_row(49 - (i % 50), args.cfg) for i in range(total)])
    cam = lambda c: PerspectiveCameras(c["R"], c["T"], c["f"], c["p"], device=dev)
    stream = current_stream(dev)
    use_graph = not args.no_graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(sharded, model=model, n=n, S=S, rows=rows, de=de, dn=dn, K=K, min_ms=2000.0):
        """W warm-up + K timed denoising steps of one scene per rank (replicas) or of one scene over all ranks (sharded);
        returns (max-over-ranks ms for the K steps, plan, scene, cameras)."""
        if sharded:
            model.shard_views()
        else:
            model.view_group = None
        sc_ = synthetic.scene_inputs(n, S, seed=0 if (sharded or world == 1) else rank)
        cams_, icams_ = cam(sc_["cams"]), cam(sc_["in_cams"])
        plan_ = model.step_plan(n, S, D, use_cfg=args.cfg != 1.0)
        model.bind_scene(plan_, cams_, sc_["input_latents"].to(dev), icams_, sc_["clip_v_embed"].to(dev), stream)
        plan_.x.copy_(sc_["x_T"].reshape(n, 5, S * S))
        plan_.set_tables(rows, de, dn)

        def one_step():
            plan_.loop_step(stream, use_graph=use_graph)  # sharded: the all-gather of the latents is the last call of the step graph

        for _ in range(max(Wm, 3)):
            one_step()
        # EXACTLY K steps per repetition between barrier + synchronize, CUDA events on the launching stream, max over ranks;
        # repeated (>= 3 times, >= 2 s of timed work) and the MEDIAN repetition is the reported one
        reps_ms, reps = [], args.reps
        x_start = plan_.x.clone()  # every repetition denoises the same latents from the same schedule position
        while True:
            plan_.counter.fill_(Wm)
            plan_.x.copy_(x_start)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(K):
                one_step()
            e1.record()
            barrier()
            ms_ = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms_], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_ = float(t)
            reps_ms.append(ms_)
            done = len(reps_ms) >= reps if reps > 0 else (len(reps_ms) >= 3 and sum(reps_ms) >= min_ms)
            if world > 1:  # one decision for all ranks
                flag = torch.tensor([1 if done else 0], device=dev)
                dist.broadcast(flag, 0)
                done = bool(flag.item())
            if done or len(reps_ms) >= 200:
                break
        ms_ = sorted(reps_ms)[len(reps_ms) // 2]
        plan_.reps_ms = reps_ms
        return ms_, plan_, sc_, cams_, icams_

    clocks = ClockSampler(local)
    other = None
    if world > 1:  # the non-headline mode first, so that the headline plan stays bound for the e2e / kernel-table legs
        o_ms, o_plan, _, _, _ = timed_loop(not shard)
        o_scenes = world if shard else 1
        other = {"value": o_scenes * K / (o_ms * 1e-3), "unit": "steps/s", "ms_per_step": o_ms / K, "views_per_gpu": o_plan.q,
                 "finite": bool(torch.isfinite(o_plan.x).all())}
    extra = {}
    if world > 1:
        # ---- the view-sharded trajectory against the single-GPU one, on hardware: 4 DDIM steps from the same x_T / noise through
        #      the public sampler, once with the views sharded over the ranks (NCCL all-gather inside the step graph), once unsharded
        sc4 = synthetic.scene_inputs(n, S, seed=0)
        de4, dn4 = synthetic.step_noises(n, D, S, 4, seed=1)
        model.ddim._make_schedule(4, "uniform", 1.0)
        outs = []
        for sharded_ in (True, False):
            if sharded_:
                model.shard_views()
            else:
                model.view_group = None
            outs.append(model.ddim.sample(cam(sc4["cams"]), sc4["input_latents"].to(dev), cam(sc4["in_cams"]), sc4["clip_v_embed"].to(dev),
                                          unconditional_scale=args.cfg, depth=True, verbose=False, x_T=sc4["x_T"], depth_eps=de4, ddim_noise=dn4))
        model.ddim._make_schedule(50, "uniform", 1.0)
        rel = ((outs[0] - outs[1]).norm() / outs[1].norm()).reshape(1)
        dist.all_reduce(rel, op=dist.ReduceOp.MAX)
        extra["rel_l2_vs_unsharded"] = float(rel)
        if other is not None and not shard:
            other["rel_l2_vs_unsharded"] = float(rel)
            other["rel_l2_note"] = "x_0 of a 4-step DDIM loop, views sharded over the ranks vs all views on one GPU (max over ranks); differs only by fp16 rounding of differently shaped GEMM batches"
    if world == 8 and os.environ.get("MVD_BENCH_EXTRA", "1") != "0" and S == 32 and n == 8:
        # ---- BASELINE configs[2] (N = 16 views, 2 views per GPU) and configs[4] (N = 8 views at 512^2 = 64x64 latents, 1 view per GPU),
        #      both view-sharded over the 8 GPUs with the in-graph all-gather; fewer steps, 3 repetitions each
        Kx = min(K, 20)
        def extra_cfg(nx, Sx, mdl):
            dex1, dnx1 = synthetic.step_noises(nx, D, Sx, Kx + Wm, seed=1)
            rowsx = torch.stack([mdl.ddim.step_row(49 - (i % 50), args.cfg) for i in range(Kx + Wm)])
            ms_x, plan_x, _, _, _ = timed_loop(True, model=mdl, n=nx, S=Sx, rows=rowsx, de=dex1, dn=dnx1, K=Kx, min_ms=0.0)
            fx = step_flops(nx, Sx, D, args.cfg)
            return {"workload": f"N={nx} views {8 * Sx}^2 view-sharded over {world} GPUs ({plan_x.q} views per GPU), one NCCL all-gather per step inside the step graph",
                    "value": Kx / (ms_x * 1e-3), "unit": "steps/s", "ms_per_step": ms_x / Kx, "steps": Kx, "gflop_per_step": round(fx / 1e9, 1),
                    "achieved_tflops_per_gpu": round(fx * Kx / (ms_x * 1e-3) / world / 1e12, 2), "finite": bool(torch.isfinite(plan_x.x).all())}
        extra["configs2_n16_sharded"] = extra_cfg(16, 32, model)
        model64 = build_model(320, 8, D=D, S=64, device=dev)
        extra["configs4_s64_sharded"] = extra_cfg(8, 64, model64)
        del model64
        torch.cuda.empty_cache()
    ms, plan, sc, cams, icams = timed_loop(shard)
    finite = bool(torch.isfinite(plan.x).all())
    scenes = 1 if (shard or world == 1) else world
    steps_per_s = scenes * K / (ms * 1e-3)

    # ---- end to end through the public sampler API with HOST buffers (pinned): per step H2D of the schedule row and the
    #      step's noise draws, D2H of the step's x_t
    model.ddim._make_schedule(50, "uniform", 1.0)
    K2 = min(K, 50)
    pin = lambda t: t.contiguous().pin_memory()
    host = dict(x_T=pin(sc["x_T"]), de=pin(de1), dn=pin(dn1))
    if K2 != 50:
        model.ddim._make_schedule([s for s in (50, 25, 20, 10, 8, 5, 4, 2, 1) if s <= K2][0], "uniform", 1.0)
    K2 = int(model.ddim.ddim_timesteps.shape[0])
    for _ in range(2):  # warm-up incl. graph capture of the host-streamed step
        model.ddim.sample(cams, sc["input_latents"].to(dev), icams, sc["clip_v_embed"].to(dev), unconditional_scale=args.cfg, depth=True,
                          verbose=False, x_T=host["x_T"], depth_eps=host["de"][:K2], ddim_noise=host["dn"][:K2], host_io=True)
    lat_h, clip_h = pin(sc["input_latents"]), pin(sc["clip_v_embed"])
    e2e_runs = []
    for _ in range(3):  # median of three end-to-end runs (host clock around barrier + synchronize, max over ranks)
        barrier()
        t0 = time.perf_counter()
        x_out = model.ddim.sample(cams, lat_h.to(dev, non_blocking=True), icams, clip_h.to(dev, non_blocking=True),
                                  unconditional_scale=args.cfg, depth=True, verbose=False, x_T=host["x_T"], depth_eps=host["de"][:K2],
                                  ddim_noise=host["dn"][:K2], host_io=True)
        x_host = x_out.cpu()
        barrier()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t)
        e2e_runs.append(e2e_s)
    e2e_s = sorted(e2e_runs)[1]
    clk = clocks.stop()
    q = plan.q
    h2d = 16 * 4 + n * D * S * S * 4 + n * 5 * S * S * 4
    d2h = q * 5 * S * S * 4
    e2e = {"value": scenes * K2 / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": K2, "runs": 3, "api": "DDIMSampler.sample(host_io=True): pinned-host schedule row + noise draws copied in every step, x_t read back every step"}

    # ---- per-kernel device times of one step (events), roofline of the dominant kernel
    pk = peaks()
    roof, ktab = None, None
    if rank == 0:
        plan.counter.zero_()
        ktab, tot_ms = kernel_table(plan._loop_prog, stream)
        plan.counter.zero_()
        top = ktab[0]
        traffic, traffic_src = None, None
        tensor_pct, tensor_pct_step = None, None
        for tname in ("r02_gemm_traffic_v6.json", "r02_gemm_traffic.json", "r01_gemm_traffic.json"):  # written by tools/ncu_summary.py from an ncu pass of this command
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath) and S == 32:
                tj = json.load(open(tpath))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = "committed file profiles/" + tname + " (not measured in this run): " + str(tj.get("source"))
                tensor_pct, tensor_pct_step = tj.get("tensor_pipe_active_pct"), tj.get("tensor_pipe_active_pct_of_step")
                break
        if top["tflops"]:
            flops = top["tflops"] * 1e12 * top["ms"] * 1e-3
            g_ms, g_n = kernel_graph_time(plan._loop_prog, top["kernel"], torch.cuda.current_stream())
            ach = flops / (g_ms * 1e-3) / 1e12 if g_ms else top["tflops"]
            roof = {"bound": "tensor", "kernel": top["kernel"], "achieved": round(ach, 2), "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": round(ach / pk["tflops"], 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": f"{pk['source']} (sustained bf16 cuBLAS)",
                    "launches_per_step": top["calls"], "share_of_step": top["share"],
                    "tensor_pipe_active_pct_ncu": tensor_pct, "tensor_pipe_active_pct_of_step_ncu": tensor_pct_step,
                    "avg_launch_us": round(g_ms * 1e3 / g_n, 2) if g_ms else None,
                    "achieved_eager_events": top["tflops"],
                    "note": "achieved = sum of 2MNK over the kernel's launches of one step / device time of exactly those launches replayed "
                            "back to back from a CUDA graph (CUDA events around the replay, best of 5; fp16 operands, fp32 accumulate); "
                            "achieved_eager_events = the same flops / sum of per-launch event brackets on the eager path (each bracket "
                            "carries a few us of launch + event overhead); share_of_step comes from the eager table"}
        if args.kernel_table:
            json.dump({"step_ms_sum_of_kernels": tot_ms, "kernels": ktab}, open(args.kernel_table, "w"), indent=1)
    if world > 1:
        dist.barrier()

    if rank == 0:
        f_step = step_flops(n, S, D, args.cfg)
        line = {"metric": METRIC, "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms / K, "timed_repetitions": len(plan.reps_ms), "ms_per_step_min_max": [round(min(plan.reps_ms) / K, 4), round(max(plan.reps_ms) / K, 4)],
                "higher_is_better": True, "scaling": "strong" if (shard or world == 1) else "weak",
                "vs_baseline": None, "dtype": "f16 operands / f32 accumulate + f32 residual stream", "data": "synthetic",
                "config": {"workload": workload_string(args),
                           "mode": ("view-sharded, 1 all-gather/step" if shard else ("replicas" if world > 1 else "single GPU")),
                           "views_per_gpu": q, "cuda_graph": use_graph,
                           "l2": "no flush: every step streams 2.08 GB of fp16 weights (>> 126 MB L2)"},
                "e2e": e2e, "gpu_launches": int((plan.kernels_per_step or 0) * K), "kernels_per_step": plan.kernels_per_step,
                "clocks": clk, "roofline": roof,
                "step_roofline": {"gflop_per_step": round(f_step / 1e9, 1), "achieved_tflops_per_gpu": round(f_step * steps_per_s / world / 1e12, 2),
                                  "frac_of_sustained_peak": round(f_step * steps_per_s / world / 1e12 / pk["tflops"], 4)},
                "kernels": ktab[:8] if ktab else None, "finite": finite and bool(torch.isfinite(x_host).all())}
        if other is not None:
            if shard:
                other["mode"] = "replicas: one independent scene per GPU, no collective (weak scaling)"
                line["replicas"] = other
            else:
                other["mode"] = "ONE scene view-sharded over the GPUs, 1 NCCL all-gather of the 5-channel latents per step, captured inside the step's CUDA graph (strong scaling)"
                line["sharded"] = other
        line.update({k: v for k, v in extra.items() if k != "rel_l2_vs_unsharded" or shard})
        if not args.no_cpu_baseline and world == 1:
            del model, plan
            torch.cuda.empty_cache()
            cb, _, _ = cpu_arm(args, steps=3, warmup=1, budget_s=60.0)
            line["cpu_baseline"] = cb
        emit(line)
    if world > 1:
        # the step graphs hold captured NCCL kernels: tearing the communicator down under them can block, so every rank leaves
        # through a last barrier + a hard exit instead of destroy_process_group()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


# ------------------------------------------------------------------------------------------------ training mode (configs[3])
def run_train(args):
    """BASELINE configs[3] (train.py forward + backward, N = 8 views, D = 3 — configs/mvd_train.yaml:28 — DDP over the GPUs, one
    scene per rank per step as train.py:55 asserts; "batch = 4 scenes" = 4 such steps): ViewFusion.forward -> loss.backward() ->
    AdamW.step().  The contractions (forward, dgrad, wgrad) run on the library's tcgen05 GEMM (mvdfusion_b200/training.py); gradients
    are all-reduced by torch DDP with bf16 compression.  Prints its own JSON line (metric: training scenes/sec)."""
    import torch.distributed as dist
    from common import build_model, standin_clip_encode, standin_vae_encode, synthetic_dataset_batch
    from mvdfusion_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, S, D, K, Wm = args.views, args.latent, 3, args.steps, max(args.warmup, 1)
    model = build_model(320, 8, D=D, S=S, device=dev)
    model.finetune_unet = True
    for p_ in model.parameters():
        p_.requires_grad_(True)
    for p_ in model.view_attn.t_embedder.parameters():   # never read by forward (find_unused_parameters=True in train.py:38)
        p_.requires_grad_(False)
    n_train = sum(p_.numel() for p_ in model.parameters() if p_.requires_grad)
    net = model
    if world > 1:
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
        net.register_comm_hook(state=None, hook=default_hooks.bf16_compress_hook)
    use_graph = bool(getattr(args, "train_graph", False)) and world == 1
    opt = torch.optim.AdamW([p_ for p_ in model.parameters() if p_.requires_grad], lr=1e-5, capturable=use_graph)
    tc = {"input_batch_size": 1, "train_batch_size": n, "random_views": False}
    batch = synthetic_dataset_batch(n + 1, 8 * S, seed=rank)
    images = batch.pop("images")
    batch["latents"] = standin_vae_encode(images, model.z_scale_factor) * 4.0
    batch["clip_embed"] = standin_clip_encode(images)
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    model.train()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = net(batch, tc)
        loss.backward()
        opt.step()
        return loss

    launches_per_replay = None
    if use_graph:
        # whole-step capture (the PyTorch whole-network recipe): eager warm-up on a side stream, gradients allocated inside the capture,
        # AdamW(capturable=True); every replay draws fresh t / noise / depth jitter from the graph-registered generator
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        c_cap = _lib.launch_count()
        with torch.cuda.graph(graph):
            static_loss = net(batch, tc)
            static_loss.backward()
            opt.step()
        launches_per_replay = _lib.launch_count() - c_cap

        def step():  # noqa: F811
            graph.replay()
            return static_loss

    for _ in range(Wm):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    c0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    launches = _lib.launch_count() - c0 if launches_per_replay is None else launches_per_replay * K
    if rank == 0:
        f_fwd = (F_GRID_GFLOP * n * n * S * S * D + n * 235.8) * 1e9   # F_U(32, D=3) = 235.8 GFLOP per view
        emit({"metric": "training scenes/sec (forward + backward + AdamW, one scene of N views per rank per step)", "value": world * K / (ms * 1e-3),
              "unit": "scenes/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f16 operands / f32 accumulate, f32 master weights + gradients, bf16-compressed all-reduce", "data": "synthetic",
              "config": {"workload": f"BASELINE configs[3]: train fwd+bwd, N={n} views {8 * S}^2, D=3, all {n_train / 1e6:.1f} M parameters trainable, DDP x{world}",
                         "native": ("contractions (forward, dgrad, wgrad) on mvd_gemm_f16; norms / activations via ATen (MVD_TRAIN_ATEN_POINTWISE=1); attention cores / gather via ATen"
                                    if os.environ.get("MVD_TRAIN_ATEN_POINTWISE") == "1" else
                                    "contractions (forward, dgrad, wgrad) on mvd_gemm_f16; LayerNorm / GroupNorm+SiLU / GELU / SiLU / GEGLU forward + backward on csrc/train.cu (ABI 15); attention cores / gather via ATen")},
              "cuda_graph": use_graph, "gpu_launches": int(launches), "loss": float(loss), "finite": bool(torch.isfinite(loss)),
              "achieved_tflops_per_gpu": round(3 * f_fwd * K / (ms * 1e-3) / 1e12, 2)})
    if world > 1:
        dist.barrier()
        os._exit(0)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints there (NCCL's version banner...) was
    re-routed to stderr at start-up."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


if __name__ == "__main__":
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.mode == "train":
        run_train(a)
    else:
        run_native(a)
