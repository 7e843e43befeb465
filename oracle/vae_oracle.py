"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's VAE decode, the step right after the
denoising loop (`demo.py:92-94` -> `ViewFusion.decode`, mvdfusion/viewfusion_zero_depth_rgb.py:162-163).

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this module; the product never does.

Follows, keyed by the reference's state-dict names (prefix e.g. "vae."):
  AutoencoderKL.decode        external/sd1/ldm/models/autoencoder.py:331-334   post_quant_conv (1x1) -> Decoder
  AutoencoderKL.encode        autoencoder.py:325-329                           Encoder -> quant_conv (1x1) -> posterior (mean | logvar)
  Decoder.forward             external/sd1/ldm/modules/diffusionmodules/model.py:541-577
  Encoder.forward             model.py:440-460;  Downsample.forward model.py:72-79 (asymmetric zero padding)
  ResnetBlock.forward         model.py:122-141   (temb is None in the decoder: temb_ch = 0)
  AttnBlock.forward           model.py:176-202   (single head, softmax(q k^T c^-1/2))
  Upsample.forward            model.py:54-58     (nearest x2, then conv3x3)
  Normalize                   model.py:38-39     (GroupNorm 32 groups, eps 1e-6)
The reference's Decoder ends with a quirk that is reproduced literally: the normalised activation is ROUNDED TO FP16
(`h_fake = self.norm_out(h).type(torch.float16)`, `h = h + (h_fake - h).detach()`, model.py:563-569) before swish + conv_out.

Pinned against the reference's own Decoder by tests/golden/make_golden_vae.py (fixture tests/golden/vae_decoder_outputs.pt).
"""
import torch
import torch.nn.functional as F


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _conv(sd, p, x, pad):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def resnet_block(sd, p, x):
    """model.py:122-141 with temb = None"""
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)), 1)
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)), 1)
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(sd, p + ".nin_shortcut", x, 0)
    return x + h


def attn_block(sd, p, x):
    """model.py:176-202"""
    b, c, hh, ww = x.shape
    h = _gn(sd, p + ".norm", x)
    q = _conv(sd, p + ".q", h, 0).reshape(b, c, hh * ww).permute(0, 2, 1)  # b, hw, c
    k = _conv(sd, p + ".k", h, 0).reshape(b, c, hh * ww)                     # b, c, hw
    v = _conv(sd, p + ".v", h, 0).reshape(b, c, hh * ww)
    w = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)             # b, hw(q), hw(k)
    o = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", o, 0)


def decoder_forward(sd, z, ch_mult, num_res_blocks, prefix="decoder"):
    """Decoder.forward (model.py:541-577); attn_resolutions = [] as in every config of the reference (configs/*.yaml)"""
    p = prefix
    h = _conv(sd, p + ".conv_in", z, 1)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    for lvl in reversed(range(len(ch_mult))):
        for j in range(num_res_blocks + 1):
            h = resnet_block(sd, f"{p}.up.{lvl}.block.{j}", h)
        if lvl != 0:
            h = _conv(sd, f"{p}.up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"), 1)
    h = _gn(sd, p + ".norm_out", h).to(torch.float16).float()  # the fp16 round trip of model.py:563-569
    return _conv(sd, p + ".conv_out", _swish(h), 1)


def encoder_forward(sd, x, ch_mult, num_res_blocks, prefix="encoder"):
    """Encoder.forward (model.py:440-460); Downsample = F.pad(x, (0,1,0,1)) + conv3x3 stride 2 padding 0 (model.py:65-76)"""
    p = prefix
    h = _conv(sd, p + ".conv_in", x, 1)
    for lvl in range(len(ch_mult)):
        for j in range(num_res_blocks):
            h = resnet_block(sd, f"{p}.down.{lvl}.block.{j}", h)
        if lvl != len(ch_mult) - 1:
            q = f"{p}.down.{lvl}.downsample.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1), mode="constant", value=0), sd[q + ".weight"], sd[q + ".bias"], stride=2, padding=0)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    return _conv(sd, p + ".conv_out", _swish(_gn(sd, p + ".norm_out", h)), 1)


def vae_encode_moments(sd, x, ch_mult=(1, 2, 4, 4), num_res_blocks=2, prefix=""):
    """AutoencoderKL.encode up to the posterior's parameters (autoencoder.py:325-329): (n, 3, R, R) in [-1, 1] -> (n, 8, R/8, R/8)
    = mean | logvar; DiagonalGaussianDistribution.mode() is the mean half (distributions.py:24-58)"""
    h = encoder_forward(sd, x, ch_mult, num_res_blocks, prefix=prefix + "encoder")
    return F.conv2d(h, sd[prefix + "quant_conv.weight"], sd[prefix + "quant_conv.bias"])


def viewfusion_encode(sd, img, z_scale_factor=0.18215, **kw):
    """ViewFusion.encode (viewfusion_zero_depth_rgb.py:158-159): vae.encode(normalize(x)).mode() * scale with
    normalize(x) = clip(2x - 1, -1, 1) (utils/common_utils.py:60-64)"""
    mom = vae_encode_moments(sd, torch.clip(img * 2 - 1.0, -1.0, 1.0), **kw)
    return mom[:, : mom.shape[1] // 2] * z_scale_factor


def vae_decode(sd, z, ch_mult=(1, 2, 4, 4), num_res_blocks=2, prefix=""):
    """AutoencoderKL.decode (autoencoder.py:331-334): z (n, 4, S, S) -> image (n, 3, 8S, 8S), roughly in [-1, 1]"""
    pre = prefix
    z = F.conv2d(z, sd[pre + "post_quant_conv.weight"], sd[pre + "post_quant_conv.bias"])
    return decoder_forward(sd, z, ch_mult, num_res_blocks, prefix=pre + "decoder")


def viewfusion_decode(sd, z, z_scale_factor=0.18215, **kw):
    """ViewFusion.decode (viewfusion_zero_depth_rgb.py:162-163): unnormalize(vae.decode(z * 1 / scale)).clip(0, 1) with
    unnormalize(x) = clip((x + 1) / 2, 0, 1) (utils/common_utils.py:66-70)"""
    return torch.clip((vae_decode(sd, z * 1 / z_scale_factor, **kw) + 1.0) / 2.0, 0.0, 1.0)
