"""The VAE decode that follows the denoising loop: `ViewFusion.decode` ->
`AutoencoderKL.decode` (external/sd1/ldm/models/autoencoder.py:331-334) -> `Decoder.forward`
(external/sd1/ldm/modules/diffusionmodules/model.py:541-577).

Same drop-in rules as the rest of the mirror: class name, constructor arguments and parameter names are the reference's
(`decoder.conv_in`, `decoder.mid.block_1.norm1`, `decoder.up.3.block.0.nin_shortcut`, `decoder.up.2.upsample.conv`,
`decoder.norm_out`, `post_quant_conv`, ...: tests/golden/make_golden_vae.py loads this module's state dict into the reference's
Decoder with strict=True), the yaml target `external.sd1.ldm.models.autoencoder.AutoencoderKL` resolves here
(mvdfusion_b200/config.py), and `decode` runs as one program of C-ABI kernel calls (engine.emit_vae_decoder) — there is no
CPU path.  `encode` (Encoder -> quant_conv -> DiagonalGaussianDistribution, autoencoder.py:325-329) runs the same way; the
training side (losses, `forward` with posterior sampling, optimizers) is not part of this tier.
"""
import torch
import torch.nn as nn

from .. import engine as E
from .sd_modules import NativeModule


def _norm(c):
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)  # model.py:38-39


class ResnetBlock(nn.Module):
    """model.py:82-141 (parameter holder; the decoder builds it with temb_channels = 0, so there is no temb_proj)"""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=0):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        if conv_shortcut or temb_channels:
            raise NotImplementedError("decoder ResnetBlocks: 1x1 shortcut, no timestep embedding")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = _norm(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = _norm(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class AttnBlock(nn.Module):
    """model.py:150-202 (parameter holder): single-head attention over the pixels of one image"""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = _norm(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)


class Upsample(nn.Module):
    """model.py:42-58 (parameter holder): nearest x2 then conv3x3"""

    def __init__(self, in_channels, with_conv=True):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("resamp_with_conv=True only")
        self.with_conv = with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)


class Downsample(nn.Module):
    """model.py:60-79 (parameter holder): zero-pad right / bottom by one, then conv3x3 stride 2 padding 0"""

    def __init__(self, in_channels, with_conv=True):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("resamp_with_conv=True only")
        self.with_conv = with_conv
        self.conv = nn.Conv2d(in_channels, in_channels, 3, 2, 0)


class Encoder(nn.Module):
    """model.py:368-438 (parameter holder with the reference's module tree and names)"""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0, resamp_with_conv=True,
                 in_channels, resolution, z_channels, double_z=True, use_linear_attn=False, attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        if list(attn_resolutions) or use_linear_attn or attn_type != "vanilla" or not double_z:
            raise NotImplementedError("encoder variant outside configs/*.yaml")
        self.ch, self.ch_mult, self.in_channels = ch, tuple(ch_mult), in_channels
        self.num_resolutions, self.num_res_blocks, self.resolution, self.z_channels = len(self.ch_mult), num_res_blocks, resolution, z_channels
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        self.down = nn.ModuleList()
        block_in = ch
        for lvl in range(self.num_resolutions):
            down = nn.Module()
            down.block, down.attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * self.ch_mult[lvl]
            for _ in range(num_res_blocks):
                down.block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout))
                block_in = block_out
            if lvl != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels, 3, 1, 1)


class DiagonalGaussianDistribution:
    """external/sd1/ldm/modules/distributions/distributions.py:24-58: the posterior over latents; parameters = mean | logvar"""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self):
        return self.mean + self.std * torch.randn(self.mean.shape).to(device=self.parameters.device)

    def mode(self):
        return self.mean


class Decoder(nn.Module):
    """model.py:462-539 (parameter holder with the reference's module tree and names)"""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0, resamp_with_conv=True,
                 in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False, use_linear_attn=False,
                 attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if list(attn_resolutions) or give_pre_end or tanh_out or use_linear_attn or attn_type != "vanilla":
            raise NotImplementedError("decoder variant outside configs/*.yaml (attention only in the middle block)")
        self.ch, self.out_ch, self.ch_mult = ch, out_ch, tuple(ch_mult)
        self.num_resolutions = len(self.ch_mult)
        self.num_res_blocks, self.resolution, self.z_channels = num_res_blocks, resolution, z_channels
        block_in = ch * self.ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.up = nn.ModuleList()
        for lvl in reversed(range(self.num_resolutions)):
            up = nn.Module()
            up.block, up.attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * self.ch_mult[lvl]
            for _ in range(num_res_blocks + 1):
                up.block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout))
                block_in = block_out
            if lvl != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
            self.up.insert(0, up)  # index = resolution level, as in the reference
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)


class AutoencoderKL(NativeModule):
    """external/sd1/ldm/models/autoencoder.py:285-334 (inference side: encode / decode)"""

    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None):
        super().__init__()
        if ckpt_path is not None:
            raise NotImplementedError("load weights through config.load_model_from_config (ViewFusion does, viewfusion_zero_depth_rgb.py:75)")
        self.image_key, self.embed_dim = image_key, embed_dim
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)

    def encode(self, x):
        """x (n, 3, R, R) fp32 in [-1, 1] on a CUDA device -> the posterior over (n, 4, R/8, R/8) latents (`.mode()`, `.sample()`)"""
        n, c, R, R2 = x.shape
        e = self.encoder
        if R != R2 or c != e.in_channels or c > 16 or (R & (R - 1)) != 0 or R < 2 ** (e.num_resolutions - 1) * 2:
            raise NotImplementedError("encode: square power-of-two images with the configured channel count")

        def make(plan, b):
            src = b.ops.empty((n, c, R * R), torch.float32)
            plan.inputs["x"] = src
            rows, S = E.emit_vae_encoder(b, src, n, R, e.ch, e.ch_mult, e.num_res_blocks, c, e.z_channels, self.embed_dim)
            dst = b.ops.empty((n, 2 * self.embed_dim, S * S), torch.float32)
            b.prog.append(b.ops.rows_to_nchw(rows, dst, n, 2 * self.embed_dim, 2 * self.embed_dim, S * S))
            plan.outputs["y"] = dst
            plan.side = S

        plan = self._plan(("encode", n, R), make)
        return DiagonalGaussianDistribution(self._execute(plan, {"x": x}).reshape(n, 2 * self.embed_dim, plan.side, plan.side))

    def decode(self, z):
        """z (n, 4, S, S) fp32 on a CUDA device -> (n, 3, 8S, 8S) fp32"""
        n, zc, S, S2 = z.shape
        d = self.decoder
        if S != S2 or zc != d.z_channels or zc > 16 or (S & (S - 1)) != 0:
            raise NotImplementedError("decode: square power-of-two latent maps with the configured channel count")

        def make(plan, b):
            src = b.ops.empty((n, zc, S * S), torch.float32)
            plan.inputs["z"] = src
            rows, H = E.emit_vae_decoder(b, src, n, S, d.ch, d.ch_mult, d.num_res_blocks, zc, d.out_ch)
            dst = b.ops.empty((n, d.out_ch, H * H), torch.float32)
            b.prog.append(b.ops.rows_to_nchw(rows, dst, n, d.out_ch, 4, H * H))
            plan.outputs["y"] = dst
            plan.side = H

        plan = self._plan(("decode", n, S), make)
        return self._execute(plan, {"z": z}).reshape(n, d.out_ch, plan.side, plan.side)

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("training forward (encode -> sample -> decode) is outside this tier")
