"""Oracle: functional fp32 restatement of ``AutoencoderKL.decode`` and ``AutoencoderKL.encode``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference: diffusers/models/vae.py:147-224
(Decoder), :575-610 (post_quant_conv + decode), unet_2d_blocks.py:320-396 (UNetMidBlock2D),
:1646-1697 (UpDecoderBlock2D), resnet.py:367-495 (ResnetBlock2D, temb=None), :77-143
(Upsample2D), attention.py:247-379 (AttentionBlock, single head, fp32 softmax).
Encoder half (the first-frame conditioning prep of scripts/inference.py:340-365):
vae.py:67-144 (Encoder), :565-573 (encode + quant_conv), :341-361 (DiagonalGaussianDistribution),
unet_2d_blocks.py DownEncoderBlock2D, resnet.py:146-188 (Downsample2D with padding=0:
F.pad(x, (0, 1, 0, 1)) then a valid stride-2 conv).
"""
import torch
import torch.nn.functional as F


def default_vae_config(**over):
    """SD-1.5 KL-f8 autoencoder."""
    cfg = dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4,
               out_channels=3, norm_num_groups=32)
    cfg.update(over)
    return cfg


def _conv(sd, p, x, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=padding)


def resnet_block_2d(sd, p, x, groups):
    """diffusers/models/resnet.py:451-495 with temb=None, eps=1e-6, output_scale_factor=1."""
    h = F.silu(F.group_norm(x, groups, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-6))
    h = _conv(sd, p + ".conv1", h)
    h = F.silu(F.group_norm(h, groups, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-6))
    h = _conv(sd, p + ".conv2", h)
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def attention_block(sd, p, x, groups):
    """diffusers/models/attention.py:331-379: one head of width C, scale C^-1/2, softmax in fp32."""
    b, c, h, w = x.shape
    y = F.group_norm(x, groups, sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], 1e-6)
    t = y.reshape(b, c, h * w).transpose(1, 2)
    q = F.linear(t, sd[p + ".query.weight"], sd[p + ".query.bias"])
    k = F.linear(t, sd[p + ".key.weight"], sd[p + ".key.bias"])
    v = F.linear(t, sd[p + ".value.weight"], sd[p + ".value.bias"])
    s = torch.matmul(q, k.transpose(1, 2)) * (c ** -0.5)
    o = torch.matmul(torch.softmax(s.float(), dim=-1), v)
    o = F.linear(o, sd[p + ".proj_attn.weight"], sd[p + ".proj_attn.bias"])
    return o.transpose(1, 2).reshape(b, c, h, w) + x


def vae_decode(sd, cfg, z):
    """AutoencoderKL.decode(z).sample for z of shape (n, 4, h, w) -> (n, 3, 8h, 8w)."""
    sd = {k: v.float() for k, v in sd.items()}
    g = cfg["norm_num_groups"]
    boc = cfg["block_out_channels"]
    x = F.conv2d(z.float(), sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])      # vae.py:576
    x = _conv(sd, "decoder.conv_in", x)
    x = resnet_block_2d(sd, "decoder.mid_block.resnets.0", x, g)
    x = attention_block(sd, "decoder.mid_block.attentions.0", x, g)
    x = resnet_block_2d(sd, "decoder.mid_block.resnets.1", x, g)
    for i in range(len(boc)):
        for j in range(cfg["layers_per_block"] + 1):
            x = resnet_block_2d(sd, f"decoder.up_blocks.{i}.resnets.{j}", x, g)
        if i < len(boc) - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(F.group_norm(x, g, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], 1e-6))
    return _conv(sd, "decoder.conv_out", x)


def vae_encode_moments(sd, cfg, x):
    """AutoencoderKL.encode(x).latent_dist.parameters (vae.py:565-568): x (n, 3, H, W) -> moments (n, 8, H/8, W/8)."""
    sd = {k: v.float() for k, v in sd.items()}
    g = cfg["norm_num_groups"]
    boc = cfg["block_out_channels"]
    h = _conv(sd, "encoder.conv_in", x.float())
    for i in range(len(boc)):
        for j in range(cfg["layers_per_block"]):
            h = resnet_block_2d(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, g)
        if i < len(boc) - 1:                                   # resnet.py:183-188: pad bottom/right by one, valid conv, stride 2
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[p + ".weight"], sd[p + ".bias"], stride=2)
    h = resnet_block_2d(sd, "encoder.mid_block.resnets.0", h, g)
    h = attention_block(sd, "encoder.mid_block.attentions.0", h, g)
    h = resnet_block_2d(sd, "encoder.mid_block.resnets.1", h, g)
    h = F.silu(F.group_norm(h, g, sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"], 1e-6))
    h = _conv(sd, "encoder.conv_out", h)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])                    # vae.py:567


def gaussian_sample(moments, noise):
    """DiagonalGaussianDistribution(moments).sample() with the normal draw supplied (vae.py:341-361):
    mean + exp(0.5 * clamp(logvar, -30, 20)) * noise."""
    mean, logvar = torch.chunk(moments.float(), 2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise


def decode_latents(sd, cfg, latents):
    """animatediff/pipelines/pipeline_animation.py:400-413: per-frame decode, (x/2+.5).clamp(0,1).
    latents (b, 4, f, h, w) -> video (b, 3, f, 8h, 8w) fp32."""
    b, c, f, h, w = latents.shape
    z = (latents.float() / 0.18215).permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    frames = torch.cat([vae_decode(sd, cfg, z[i:i + 1]) for i in range(b * f)])
    video = frames.reshape(b, f, *frames.shape[1:]).permute(0, 2, 1, 3, 4)
    return (video / 2 + 0.5).clamp(0, 1)
