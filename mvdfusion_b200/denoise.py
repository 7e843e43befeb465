"""Step plan of the multi-view denoising loop: buffers + programs for one (N views, S, D, cfg, view-shard) shape.

Reference path covered: DDIMSampler.sample loop body (mvdfusion/sampler.py:119-142) ->
ViewFusion.apply_model (mvdfusion/viewfusion_zero_depth_rgb.py:282-345) -> GridAttn.forward
(mvdfusion/view_attn_efficient2.py:413-442) -> UNetWrapper.forward / predict_with_unconditional_scale
(mvdfusion/unet.py:129-196) -> denoise_apply_impl (mvdfusion/sampler.py:42-66).

Differences from the reference's schedule of work (results are the same):
  * the conditional and unconditional UNet passes run as ONE batch of 2N images (the reference runs them back to
    back, unet.py:192-193); the unconditional half sees zero concat channels, a zero frustum pyramid and the
    per-layer CLIP vector reduced to `to_out.bias` (to_v(0) == 0);
  * loop-invariant work (cc_projection, the 16 one-token CLIP cross-attentions) runs once per scene;
  * per-step scalars (t, alpha tables, sigma) and the step's pre-drawn noise are selected on the device from
    tables by a device-resident step counter, so the whole step is one CUDA-graph replay with no host traffic.
"""
import os

import torch

from . import engine as E

STEPC_LEN = 16  # per-step constants row: [t, sqrt_acp[t], depth_std, -, a_t, a_prev, sqrt(1-a_t), sigma, add_noise, cfg, ...]


def pack_cameras(R, T, f, p):
    """[n, 16] = R (9, row-major) | T (3) | focal (2) | principal point (2)   (pytorch3d PerspectiveCameras fields)"""
    n = R.shape[0]
    return torch.cat([R.reshape(n, 9), T.reshape(n, 3), f.reshape(n, 2), p.reshape(n, 2)], dim=1).float().contiguous()


class StepPlan:
    def __init__(self, ops, state_dict, unet_spec, *, n_views, S, D, use_cfg, q_first=0, q_count=None, num_layers=3,
                 grid_heads=8, depth_scale=2.0, depth_shift=0.5, cond_batched=False, use_depth_override=False,
                 use_cond_scale=False, prefixes=("view_attn.", "unet_model.unet_model.", "")):
        self.ops = ops
        self.spec = unet_spec
        self.N, self.S, self.D = n_views, S, D
        self.hw = S * S
        self.q_first = q_first
        self.q = q_count if q_count is not None else n_views
        self.use_cfg = use_cfg
        self.n_unet = self.q * (2 if use_cfg else 1)
        self.cond_batched = cond_batched
        dev = ops.device
        N, q, hw, n_unet = self.N, self.q, self.hw, self.n_unet
        self.W_grid = E.PackedWeights(state_dict, ops, prefixes[0])
        self.W_unet = E.PackedWeights(state_dict, ops, prefixes[1])
        self.W_top = E.PackedWeights(state_dict, ops, prefixes[2])

        z32 = lambda *s: ops.zeros(s, torch.float32)
        # ---- persistent inputs / state
        self.x = z32(N, 5, hw)                      # x_t of ALL views (NCHW); rows [q_first, q_first+q) are this rank's
        self.input_latent = z32(N if cond_batched else 1, 5, hw)
        self.cams = z32(N + 1, 16)
        self.mask = ops.zeros((N,), torch.float32) + 1.0
        self.stepc = z32(STEPC_LEN)
        self.depth_eps = z32(N, D, hw)
        self.ddim_noise = z32(N, 5, hw)
        self.depth_override = z32(N, hw) if use_depth_override else None
        self.cond_scale = (z32(q) + 1.0) if use_cond_scale else None
        self.clip_in = z32(q, 796)
        self.clip_ctx = z32(n_unet, E.CTX_DIM)      # cc_projection output; unconditional rows stay zero
        self.eps_out = z32(q, 5, hw)
        self.x0_out = z32(q, 5, hw)
        self.counter = ops.zeros((1,), torch.int32)
        self.t_dev, self.scal, self.coef = self.stepc[0:1], self.stepc[1:3], self.stepc[4:10]
        self.x_local = self.x[q_first:q_first + q]
        self.noise_local = self.ddim_noise[q_first:q_first + q]
        self.stem_hilo = int(os.environ.get("MVD_HILO", "2")) >= 1
        self.c_in_pad = 32 if self.stem_hilo else 16
        self.x_in16 = ops.zeros((n_unet * hw, self.c_in_pad), torch.float16)
        self.pyramid16 = [ops.zeros((n_unet * (S >> l) * (S >> l) * D, E.CTX_DIM), torch.float16)
                          for l in range(len(unet_spec.mult))]
        self.freqs_unet = E.timestep_freqs(unet_spec.mc, dev)
        self.freqs_grid = E.timestep_freqs(256, dev)
        self.harm_freqs = ((2.0 ** torch.arange(7, dtype=torch.float32)) * 0.1).to(dev)
        half = 1.0 / float(S)
        self.ndc_grid = torch.linspace(1.0 - half, -1.0 + half, S, dtype=torch.float32).to(dev)

        arena = E.Arena(ops)
        # ---- per-scene program: cc_projection + the one-token CLIP cross-attention vectors
        b = E.Builder(ops, self.W_top, arena=arena)
        h1 = b.t32(q, 768)
        h2 = b.t32(q, 768)
        b.prog.append(ops.gemv(self.clip_in, self.W_top.lin("cc_projection.0.weight"), self.W_top.f32("cc_projection.0.bias"),
                               h1, q, 768, 796, silu_out=True))
        b.prog.append(ops.gemv(h1, self.W_top.lin("cc_projection.2.weight"), self.W_top.f32("cc_projection.2.bias"), h2, q,
                               768, 768, silu_out=True))
        b.prog.append(ops.gemv(h2, self.W_top.lin("cc_projection.4.weight"), self.W_top.f32("cc_projection.4.bias"),
                               self.clip_ctx, q, 768, 768))
        b.free(h1, h2)
        b.W = self.W_unet
        self.clipvecs = {}
        for p, C in unet_spec.st_layers():
            self.clipvecs[p] = b.clip_vector(self.clip_ctx, p, n_unet, C)
        self.scene_prog = b.prog

        # ---- per-step core: GridAttn -> pyramid -> UNet
        b = E.Builder(ops, self.W_top, arena=arena)
        b.prog.mark("time_embed")
        c_embed = b.time_mlp(self.t_dev, self.freqs_grid, 256, "time_embed.0", "time_embed.2", 256, 256)
        b.W = self.W_grid
        E.emit_gridattn(b, noisy=self.x, input_latent=self.input_latent[:1], depth_override=self.depth_override,
                        depth_eps=self.depth_eps, scal=self.scal, cams=self.cams, mask=self.mask, c_embed=c_embed,
                        n_views=N, S=S, D=D, q_first=q_first, q_count=q, num_layers=num_layers, num_heads=grid_heads,
                        depth_scale=depth_scale, depth_shift=depth_shift, frustum_out=self.pyramid16[0],
                        harm_freqs=self.harm_freqs, ndc_grid=self.ndc_grid)
        b.prog.mark("frustum_pyramid")
        E.emit_pyramid(b, self.pyramid16, q, S, D)
        self.grid_calls = len(b.prog)
        b.W = self.W_unet
        cond = self.input_latent[q_first:q_first + q] if cond_batched else self.input_latent
        b.prog.append(ops.unet_input(self.x_local, cond, cond_batched, self.cond_scale, self.x_in16, q, n_unet, hw, self.c_in_pad,
                                     hilo=self.stem_hilo))
        self.head = E.emit_unet(b, unet_spec, self.x_in16, n_unet, S, D, self.t_dev, self.freqs_unet, self.clipvecs,
                                self.pyramid16, c_in_pad=self.c_in_pad, stem_hilo=self.stem_hilo)
        self.core_prog = b.prog

        # ---- epilogues
        self.eps_prog = E.Program()
        self.eps_prog.append(ops.cfg_ddim(self.head, 8, use_cfg, self.coef, None, None, self.eps_out, None, None, q, hw))
        self.ddim_prog = E.Program()
        self.ddim_prog.mark("cfg_ddim_update")
        self.ddim_prog.append(ops.cfg_ddim(self.head, 8, use_cfg, self.coef, self.x_local, self.noise_local, self.eps_out,
                                           self.x_local, self.x0_out, q, hw))
        self.ddim_prog.mark("end")
        self.arena_bytes = arena.total_bytes
        # view-sharded mode: the exchange of the updated latents (mvdfusion_b200.mvdfusion.viewfusion_zero_depth_rgb.ViewFusion.gather_views)
        # is the LAST call of the step program, so that it is captured into the step's CUDA graph with the kernels
        self.after_step = None
        self._tables = None
        self._loop_prog = None
        self._host_prog = None
        self._graphs = {}
        self.kernels_per_step = None

    # ------------------------------------------------------------------ scene / inputs
    def set_scene(self, cams_R, cams_T, cams_f, cams_p, in_R, in_T, in_f, in_p, input_latents, clip_v_embed, stream):
        """Upload the loop-invariant inputs of one scene and run the per-scene program."""
        cam = pack_cameras(torch.cat([cams_R, in_R[:1]]), torch.cat([cams_T, in_T[:1]]), torch.cat([cams_f, in_f[:1]]),
                           torch.cat([cams_p, in_p[:1]]))
        self.cams.copy_(cam)
        il = input_latents.reshape(input_latents.shape[0], 5, self.hw).float()
        self.input_latent.copy_(il if self.cond_batched else il[:1])
        ce = clip_v_embed.reshape(-1, clip_v_embed.shape[-1]).float()
        self.clip_in.copy_(ce[self.q_first:self.q_first + self.q] if ce.shape[0] == self.N else ce)
        if stream != "defer":
            self.scene_prog.run(stream)

    def run_eps_with_drop(self, scene_args, keep_clip, keep_vol, stream):
        """apply_model on the reference's cfg == 1.0 branch, which runs the UNet wrapper with is_train=True and therefore
        drops conditions at random (mvdfusion/viewfusion_zero_depth_rgb.py:324-331, mvdfusion/unet.py:140-151):
        keep_clip / keep_vol are the per-view {0,1} masks of the CLIP embedding and of the frustum features (the concat
        mask is `cond_scale`, applied by the input-assembly kernel).  Masking is host-side glue between program segments."""
        self.set_scene(*scene_args, "defer")
        q = self.q
        for c in self.scene_prog.calls[:3]:       # cc_projection
            c(stream)
        self.clip_ctx[:q].mul_(keep_clip.view(q, 1))
        for c in self.scene_prog.calls[3:]:       # per-layer CLIP vectors
            c(stream)
        for c in self.core_prog.calls[:self.grid_calls]:
            c(stream)
        for lvl in self.pyramid16:
            lvl.view(self.n_unet, -1)[:q].mul_(keep_vol.view(q, 1).to(lvl.dtype))
        for c in self.core_prog.calls[self.grid_calls:]:
            c(stream)
        self.eps_prog.run(stream)

    def set_step_constants(self, row):
        self.stepc.copy_(row.to(self.stepc.device, non_blocking=True))

    # ------------------------------------------------------------------ single call (ViewFusion.apply_model)
    def run_eps(self, stream):
        self.core_prog.run(stream)
        self.eps_prog.run(stream)

    # ------------------------------------------------------------------ loop (DDIMSampler.sample)
    def set_tables(self, step_rows, depth_eps_all, noise_all):
        """step_rows [steps, 16]; depth_eps_all [steps, N, D, hw]; noise_all [steps, N, 5, hw].  The device tables are
        persistent (re-allocated only when the step count grows), so the captured loop graph stays valid across scenes."""
        steps = step_rows.shape[0]
        ops = self.ops
        if self._tables is None or self._tables[0].shape[0] < steps:
            self._tables = (ops.zeros((steps, STEPC_LEN), torch.float32), ops.zeros((steps, self.N * self.D * self.hw), torch.float32),
                            ops.zeros((steps, self.N * 5 * self.hw), torch.float32))
            pro = E.Program()
            pro.append(ops.gather_rows(self._tables[0], STEPC_LEN, self.counter, self.stepc))
            pro.append(ops.gather_rows(self._tables[1], self.N * self.D * self.hw, self.counter, self.depth_eps))
            pro.append(ops.gather_rows(self._tables[2], self.N * 5 * self.hw, self.counter, self.ddim_noise))
            pro.append(ops.increment(self.counter, 1))
            loop = E.Program()
            loop.extend(pro)
            loop.extend(self.core_prog)
            loop.extend(self.ddim_prog)
            if self.after_step is not None:
                loop.append(self.after_step)
            self._loop_prog = loop
            self._graphs.pop("loop", None)
        self._tables[0][:steps].copy_(step_rows.float(), non_blocking=True)
        self._tables[1][:steps].copy_(depth_eps_all.reshape(steps, -1).float(), non_blocking=True)
        self._tables[2][:steps].copy_(noise_all.reshape(steps, -1).float(), non_blocking=True)
        self.counter.zero_()

    def _replay(self, key, prog, stream, use_graph):
        """Run `prog` once: call by call, or (CUDA) from a graph captured on first use."""
        if not use_graph or self.x.device.type != "cuda":
            prog.run(stream)
            return
        if key not in self._graphs:
            # warm-up outside capture (lazy module loading, smem attribute opt-ins), then restore the loop state
            from . import _lib
            x_keep, c_keep = self.x.clone(), self.counter.clone()
            c0 = _lib.launch_count()
            prog.run(stream)
            self.kernels_per_step = _lib.launch_count() - c0
            torch.cuda.synchronize()
            self.x.copy_(x_keep)
            self.counter.copy_(c_keep)
            g = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread of a view-sharded run must not invalidate the capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                prog.run(torch.cuda.current_stream().cuda_stream)
            self.x.copy_(x_keep)
            self.counter.copy_(c_keep)
            self._graphs[key] = g
        self._graphs[key].replay()

    def loop_step(self, stream, use_graph=True):
        """One DDIM iteration for this rank's views, everything selected on the device (advances the step counter)."""
        self._replay("loop", self._loop_prog, stream, use_graph)

    def host_step(self, stream, row, depth_eps, ddim_noise, use_graph=True):
        """One DDIM iteration whose per-step inputs come from (pinned) HOST memory: constants row [16], depth jitter
        (N,D,hw) and DDIM noise (N,5,hw) are copied host->device on `stream` order, then the step program runs."""
        self.stepc.copy_(row, non_blocking=True)
        self.depth_eps.copy_(depth_eps.reshape(self.depth_eps.shape), non_blocking=True)
        self.ddim_noise.copy_(ddim_noise.reshape(self.ddim_noise.shape), non_blocking=True)
        if self._host_prog is None:
            self._host_prog = E.Program()
            self._host_prog.extend(self.core_prog)
            self._host_prog.extend(self.ddim_prog)
            if self.after_step is not None:
                self._host_prog.append(self.after_step)
        self._replay("host", self._host_prog, stream, use_graph)


class UNetStagePlan:
    """UNetWrapper.forward / predict_with_unconditional_scale (mvdfusion/unet.py:129-196) as one program:
    frustum pyramid -> input assembly -> UNet over q (x2 under CFG) images -> CFG combine."""

    def __init__(self, ops, weights, unet_spec, *, n_views, S, D, use_cfg):
        self.ops, self.spec = ops, unet_spec
        self.q, self.S, self.D, self.hw = n_views, S, D, S * S
        self.use_cfg = use_cfg
        q, hw = self.q, self.hw
        n_unet = self.n_unet = q * (2 if use_cfg else 1)
        z32 = lambda *s: ops.zeros(s, torch.float32)
        self.x = z32(q, 5, hw)
        self.cond = z32(q, 5, hw)
        self.cond_scale = z32(q) + 1.0
        self.clip_ctx = z32(n_unet, E.CTX_DIM)
        self.vol32 = z32(q * hw * D, E.CTX_DIM)
        self.stepc = z32(STEPC_LEN)
        self.t_dev, self.coef = self.stepc[0:1], self.stepc[4:10]
        self.eps_out = z32(q, 5, hw)
        self.stem_hilo = int(os.environ.get("MVD_HILO", "2")) >= 1
        self.c_in_pad = 32 if self.stem_hilo else 16
        self.x_in16 = ops.zeros((n_unet * hw, self.c_in_pad), torch.float16)
        self.pyramid16 = [ops.zeros((n_unet * (S >> l) * (S >> l) * D, E.CTX_DIM), torch.float16)
                          for l in range(len(unet_spec.mult))]
        self.freqs = E.timestep_freqs(unet_spec.mc, ops.device)
        b = E.Builder(ops, weights)
        clipvecs = {p: b.clip_vector(self.clip_ctx, p, n_unet, C) for p, C in unet_spec.st_layers()}
        b.prog.append(ops.cast(self.vol32, self.pyramid16[0], q * hw * D * E.CTX_DIM))
        E.emit_pyramid(b, self.pyramid16, q, S, D)
        b.prog.append(ops.unet_input(self.x, self.cond, True, self.cond_scale, self.x_in16, q, n_unet, hw, self.c_in_pad, hilo=self.stem_hilo))
        self.head = E.emit_unet(b, unet_spec, self.x_in16, n_unet, S, D, self.t_dev, self.freqs, clipvecs, self.pyramid16,
                                c_in_pad=self.c_in_pad, stem_hilo=self.stem_hilo)
        b.prog.append(ops.cfg_ddim(self.head, 8, use_cfg, self.coef, None, None, self.eps_out, None, None, q, hw))
        self.prog = b.prog

    def run(self, x, t, clip_embed, volume_feats, x_concat, cfg_scale, stream, cond_scale=None):
        q, hw = self.q, self.hw
        self.x.copy_(x.reshape(q, 5, hw))
        self.cond.copy_(x_concat.reshape(q, 5, hw))
        if cond_scale is None:
            self.cond_scale.fill_(1.0)
        else:
            self.cond_scale.copy_(cond_scale)
        self.clip_ctx[:q].copy_(clip_embed.reshape(q, E.CTX_DIM))
        self.vol32.copy_(volume_feats.reshape(q * hw * self.D, E.CTX_DIM))
        row = torch.zeros(STEPC_LEN)
        row[0] = float(t)
        row[9] = float(cfg_scale)
        self.stepc.copy_(row)
        self.prog.run(stream)
        return self.eps_out.reshape(q, 5, self.S, self.S).clone()
