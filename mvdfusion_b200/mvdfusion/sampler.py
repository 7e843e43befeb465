"""mvdfusion/sampler.py of the reference: the DDIM (eta = 1 -> stochastic) sampling loop, the HOT LOOP of the path.

`sample()` keeps the reference's signature and return values but drives a StepPlan: all per-step scalars and the
pre-drawn noise live in device tables, each iteration is one CUDA-graph replay of the step program, and (when the views
of a scene are sharded over ranks) one all-gather of the updated 5-channel latents.
"""
import numpy as np
import torch

from ..denoise import STEPC_LEN
from ..runtime import current_stream


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """external/sd1/ldm/modules/diffusionmodules/util.py:46-60"""
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * 0.8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    return ddim_timesteps + 1


class DDIMSampler:
    """mvdfusion/sampler.py:13-147"""

    def __init__(self, model, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, latent_size=32,
                 overwrite_x_noisy=False, z_dim=4, feed_prev_depth=False):
        self.model = model
        self.ddpm_num_timesteps = model.scheduler.num_timesteps
        self.latent_size = latent_size
        self._make_schedule(ddim_num_steps, ddim_discretize, ddim_eta, verbose=False)
        self.eta = ddim_eta
        self.overwrite_x_noisy = overwrite_x_noisy
        self.z_dim = z_dim
        self.feed_prev_depth = feed_prev_depth

    def _make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        """mvdfusion/sampler.py:25-39 (tables computed on the host in fp64, kept as fp32 like the reference)."""
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps, verbose)
        ts = torch.from_numpy(self.ddim_timesteps.astype(np.int64))
        acp = self.model.scheduler.alphas_cumprod.detach().cpu()
        a = acp[ts].double()
        a_prev = torch.cat([acp[0:1], acp[ts[:-1]]], 0)
        sig = ddim_eta * torch.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
        self.ddim_alphas_raw = self.model.scheduler.alphas.detach().cpu()[ts].float()
        self.ddim_sigmas = sig.float()
        self.ddim_alphas = a.float()
        self.ddim_alphas_prev = a_prev.float()
        self.ddim_sqrt_one_minus_alphas = torch.sqrt(1.0 - self.ddim_alphas).float()

    def step_row(self, index, cfg_scale):
        """Per-step constants consumed by the kernels (denoise.STEPC_LEN floats)."""
        sch = self.model.scheduler
        t = int(self.ddim_timesteps[index])
        sac = float(sch.sqrt_alphas_cumprod[t])
        std = float(sch.sqrt_one_minus_alphas_cumprod[t] / sch.sqrt_alphas_cumprod[t] / 10.0)
        row = torch.zeros(STEPC_LEN)
        row[0], row[1], row[2] = float(t), sac, std
        row[4], row[5] = float(self.ddim_alphas[index]), float(self.ddim_alphas_prev[index])
        row[6], row[7] = float(self.ddim_sqrt_one_minus_alphas[index]), float(self.ddim_sigmas[index])
        row[8] = 0.0 if index == 0 else 1.0
        row[9] = float(cfg_scale)
        return row

    @torch.no_grad()
    def denoise_apply_impl(self, x_target_noisy, index, noise_pred, is_step0=False, noise=None):
        """mvdfusion/sampler.py:42-66 as a standalone call (the loop fuses it with the CFG combine)."""
        device = x_target_noisy.device
        a_t = self.ddim_alphas[index].to(device).float()
        a_prev = self.ddim_alphas_prev[index].to(device).float()
        somat = self.ddim_sqrt_one_minus_alphas[index].to(device).float()
        sigma = self.ddim_sigmas[index].to(device).float()
        pred_x0 = (x_target_noisy - somat * noise_pred) / a_t.sqrt()
        dir_xt = torch.clamp(1.0 - a_prev - sigma ** 2, min=1e-7).sqrt() * noise_pred
        x_prev = a_prev.sqrt() * pred_x0 + dir_xt
        if not is_step0:
            x_prev = x_prev + sigma * (noise if noise is not None else torch.randn_like(x_target_noisy))
        return x_prev, pred_x0

    @torch.no_grad()
    def denoise_apply(self, x_target_noisy, batch_cameras, input_latents, input_cameras, clip_embed, time_steps, index,
                      is_step0=False, prev_depth=None, cfg_scale=1.0):
        """mvdfusion/sampler.py:68-88"""
        kw = dict(prev_depth=prev_depth) if self.feed_prev_depth else {}
        noise = self.model.apply_model(x_target_noisy, batch_cameras, input_latents, input_cameras, clip_embed, time_steps,
                                       cfg_scale=cfg_scale, **kw)
        return self.denoise_apply_impl(x_target_noisy, index, noise, is_step0)

    @torch.no_grad()
    def sample(self, batch_cameras, input_latents, input_cameras, clip_embed, unconditional_scale=1.0, depth=False,
               return_intermediates=False, verbose=True, x_T=None, depth_eps=None, ddim_noise=None, use_graph=True,
               host_io=False):
        """mvdfusion/sampler.py:90-147.  Extra keyword arguments (all optional): x_T (B,5,S,S), depth_eps / ddim_noise
        (steps,B,D,S,S) / (steps,B,5,S,S) inject the random draws of iteration i (i = 0 is the largest timestep) instead of
        drawing them here; use_graph=False replays the step program call by call; host_io=True keeps the per-step inputs
        (schedule constants, noise draws) in pinned HOST memory, copies them to the device every step and reads every
        step's x_t back to the host (the end-to-end mode bench.py times)."""
        if not depth:
            raise NotImplementedError("GridAttn needs the 4+1 channel latents (view_attn_efficient2.py:440): call with depth=True")
        model = self.model
        S, C = self.latent_size, self.z_dim
        B = clip_embed.shape[0]
        device = model._device.device
        D = model.view_attn.n_pts_per_ray
        total = self.ddim_timesteps.shape[0]
        if verbose:
            print(f"unconditional scale {unconditional_scale:.1f}")
        # random draws in the reference's order: x_T, then per step the depth jitter followed by the DDIM noise
        x = x_T.to(device).float() if x_T is not None else torch.randn([B, C + 1, S, S], device=device)
        slow = self.feed_prev_depth or self.overwrite_x_noisy or (unconditional_scale == 1.0 and model.drop_conditions)
        if slow:
            return self._sample_stepwise(x, batch_cameras, input_latents, input_cameras, clip_embed, unconditional_scale,
                                         return_intermediates)
        if depth_eps is None or ddim_noise is None:
            de, dn = [], []
            for i in range(total):
                de.append(torch.randn(B, D, S, S, device=device))
                dn.append(torch.randn(B, C + 1, S, S, device=device))
            depth_eps = torch.stack(de) if depth_eps is None else depth_eps
            ddim_noise = torch.stack(dn) if ddim_noise is None else ddim_noise
        rows = torch.stack([self.step_row(total - i - 1, unconditional_scale) for i in range(total)])

        plan = model.step_plan(B, S, D, use_cfg=unconditional_scale != 1.0)
        stream = current_stream(device)
        model.bind_scene(plan, batch_cameras, input_latents, input_cameras, clip_embed, stream)
        plan.x.copy_(x.reshape(B, 5, S * S))
        inter = []
        group = model.view_group
        if group is not None and x_T is None:
            # every rank drew its own x_T: the owning rank's rows are the authoritative ones (GridAttn of step 0 must see the
            # latents the owners denoise), so exchange them once before the loop
            model.gather_views(plan)
        if host_io:
            pin = (lambda t: t.pin_memory()) if device.type == "cuda" else (lambda t: t)
            rows_h, de_h, dn_h = pin(rows.contiguous()), pin(depth_eps.float().cpu().contiguous()), pin(ddim_noise.float().cpu().contiguous())
            # the pinned read-back buffer is kept across calls (page-locking 8 MB costs milliseconds)
            shape = (total, plan.q, 5, S, S)
            cache = self.__dict__.setdefault("_pinned_xh", {})
            if shape not in cache:
                cache.clear()
                cache[shape] = pin(torch.empty(shape))
            x_h = cache[shape]
        else:
            plan.set_tables(rows, depth_eps, ddim_noise)
        copied = None
        for i in range(total):
            if host_io:
                plan.host_step(stream, rows_h[i], de_h[i], dn_h[i], use_graph=use_graph)
            else:
                plan.loop_step(stream, use_graph=use_graph)
            # (view-sharded: the all-gather of the updated latents is the last call of the step program — inside the graph)
            if host_io:
                # the step's x_t lands in pinned host memory on the stream; the host only waits for the PREVIOUS step's
                # copy, so that it has step i+1 (H2D of its inputs + graph launch) queued while step i still runs
                x_h[i].copy_(plan.x_local.reshape(plan.q, 5, S, S), non_blocking=True)
                if device.type == "cuda":
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(device))
                    if copied is not None:
                        copied.synchronize()
                    copied = ev
            if return_intermediates:
                inter.append({"t": int(self.ddim_timesteps[total - i - 1]), "xt": plan.x.reshape(B, 5, S, S).clone(),
                              "x0": plan.x0_out.reshape(-1, 5, S, S).clone()})
        if copied is not None:
            copied.synchronize()
        if host_io:
            self.host_trajectory = x_h  # (steps, views, 5, S, S): every step's x_t as it arrived in pinned host memory (reused by the next call)
        out = plan.x.reshape(B, 5, S, S).clone()
        return (out, inter) if return_intermediates else out

    def _sample_stepwise(self, x, batch_cameras, input_latents, input_cameras, clip_embed, cfg_scale, return_intermediates):
        """The loop exactly as written in the reference (one apply_model + host-side update per step); used for the
        reference's side modes (feed_prev_depth, overwrite_x_noisy, the cfg == 1.0 condition-drop quirk)."""
        total = self.ddim_timesteps.shape[0]
        B = x.shape[0]
        prev_depth, inter = None, []
        model = self.model
        sharded = model.view_group is not None
        if sharded:  # rank-local rows of the initial latents are authoritative (each rank may have drawn its own x_T)
            x = model.all_gather_rows(model.local_rows(x))
        for i, step in enumerate(np.flip(self.ddim_timesteps)):
            index = total - i - 1
            t = torch.full((B,), int(step), device=x.device, dtype=torch.long)
            if self.overwrite_x_noisy:
                x[0] = input_latents[0].clone()
            if not sharded:
                x, x0 = self.denoise_apply(x, batch_cameras, input_latents, input_cameras, clip_embed, t, index,
                                           is_step0=index == 0, prev_depth=prev_depth if self.feed_prev_depth else None,
                                           cfg_scale=cfg_scale)
            else:
                # view-sharded: apply_model returns the predicted noise of THIS rank's views only; update those rows and
                # exchange the updated latents (and x_0, which feeds prev_depth) with one all-gather each
                kw = dict(prev_depth=prev_depth) if self.feed_prev_depth else {}
                eps = model.apply_model(x, batch_cameras, input_latents, input_cameras, clip_embed, t, cfg_scale=cfg_scale, **kw)
                xl, x0l = self.denoise_apply_impl(model.local_rows(x), index, eps, is_step0=index == 0)
                x, x0 = model.all_gather_rows(xl), model.all_gather_rows(x0l)
            prev_depth = x0[:, 4:].clone()
            inter.append({"t": int(step), "xt": x, "x0": x0})
        return (x, inter) if return_intermediates else x
