"""AutoencoderKL: reference call surface (diffusers/models/vae.py:500-638), CUDA engine underneath.

The decoder half is SURVEY 8a row a12: ``decode(z).sample`` and the batched ``decode_frames`` used by
``AnimationPipeline.decode_latents``.  All frames of a clip are decoded in ONE batch (the reference loops batch-1
calls, pipeline_animation.py:405-408); GroupNorm statistics are per image so the result is identical.
The mid-block attention (one 512-wide head over H*W tokens, softmax in fp32 - diffusers/models/attention.py:331-379)
is three tensor-core GEMMs (QK^T with fp32 scores, PV on V^T) around an fp32 row-softmax kernel.
The encoder half (``encode`` -> ``latent_dist``; SURVEY 8f row 1, the first-frame conditioning prep of
scripts/inference.py:340-365) reuses the same kernels; its downsamplers are the bottom/right-padded stride-2
convolution of diffusers Downsample2D(padding=0) (``pad_mode=1`` of fyc_conv3x3).
"""
from collections import OrderedDict
from dataclasses import dataclass

import torch

from . import ops
from .modeling import FrozenDict, ParamTreeModel


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class DiagonalGaussianDistribution:
    """diffusers/models/vae.py:341-397: moments (n, 2c, h, w) -> mean / clamped logvar / std / var, ``sample`` and ``mode``.
    A per-clip one-off on a 2 x 4 x 64 x 64 tensor: plain tensor arithmetic on the device the moments live on."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device)
        return self.mean + self.std * noise.to(self.parameters.dtype)

    def mode(self):
        return self.mean

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.0])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0 - self.logvar
                               + other.logvar, dim=[1, 2, 3])


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution


def vae_param_spec(cfg):
    spec = OrderedDict()
    boc = tuple(cfg["block_out_channels"])
    lc = cfg["latent_channels"]
    lpb = cfg["layers_per_block"]

    def conv(p, o, i, k):
        spec[p + ".weight"] = (o, i, k, k); spec[p + ".bias"] = (o,)

    def norm(p, c):
        spec[p + ".weight"] = (c,); spec[p + ".bias"] = (c,)

    def lin(p, o, i):
        spec[p + ".weight"] = (o, i); spec[p + ".bias"] = (o,)

    def resnet(p, i, o):
        norm(p + ".norm1", i); conv(p + ".conv1", o, i, 3); norm(p + ".norm2", o); conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".conv_shortcut", o, i, 1)

    def mid(p, c):
        for nm in ("group_norm",):
            norm(p + ".attentions.0." + nm, c)
        for nm in ("query", "key", "value", "proj_attn"):
            lin(p + ".attentions.0." + nm, c, c)
        resnet(p + ".resnets.0", c, c); resnet(p + ".resnets.1", c, c)

    conv("encoder.conv_in", boc[0], cfg["in_channels"], 3)
    out_c = boc[0]
    for i in range(len(boc)):
        in_c, out_c = out_c, boc[i]
        for j in range(lpb):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", in_c if j == 0 else out_c, out_c)
        if i < len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", out_c, out_c, 3)
    mid("encoder.mid_block", boc[-1])
    norm("encoder.conv_norm_out", boc[-1]); conv("encoder.conv_out", 2 * lc, boc[-1], 3)
    conv("decoder.conv_in", boc[-1], lc, 3)
    rev = boc[::-1]
    out_c = rev[0]
    for i in range(len(boc)):
        prev, out_c = out_c, rev[i]
        for j in range(lpb + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else out_c, out_c)
        if i < len(boc) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", out_c, out_c, 3)
    mid("decoder.mid_block", boc[-1])
    norm("decoder.conv_norm_out", boc[0]); conv("decoder.conv_out", cfg["out_channels"], boc[0], 3)
    conv("quant_conv", 2 * lc, 2 * lc, 1); conv("post_quant_conv", lc, lc, 1)
    return spec


class AutoencoderKL(ParamTreeModel):
    def __init__(self, in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",),
                 up_block_types=("UpDecoderBlock2D",), block_out_channels=(64,), layers_per_block=1, act_fn="silu",
                 latent_channels=4, norm_num_groups=32, sample_size=32, **unused):
        super().__init__()
        kw = dict(in_channels=in_channels, out_channels=out_channels, down_block_types=tuple(down_block_types),
                  up_block_types=tuple(up_block_types), block_out_channels=tuple(block_out_channels),
                  layers_per_block=layers_per_block, act_fn=act_fn, latent_channels=latent_channels,
                  norm_num_groups=norm_num_groups, sample_size=sample_size)
        if act_fn != "silu":
            raise NotImplementedError(act_fn)
        self.config = FrozenDict(dict(kw, _class_name="AutoencoderKL", _diffusers_version="0.11.1"))
        self._cfg = kw
        self.use_slicing = False
        self._build_tree(vae_param_spec(kw))

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = {k: v for k, v in dict(config).items() if not k.startswith("_")}
        cfg.update(kwargs)
        return cls(**cfg)

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **kw):
        import json, os
        if subfolder is not None:
            path = os.path.join(path, subfolder)
        with open(os.path.join(path, "config.json")) as f:
            model = cls.from_config(json.load(f))
        sd_file = os.path.join(path, "diffusion_pytorch_model.bin")
        model.load_state_dict(torch.load(sd_file, map_location="cpu"), strict=False)
        return model

    def enable_slicing(self):
        self.use_slicing = True      # frames are independent either way; kept for API compatibility

    def disable_slicing(self):
        self.use_slicing = False

    def encode_nhwc(self, x):
        """x [N, H, W, 3] channels-last in the compute dtype -> moments [N, H/8, W/8, 2 * latent] (compute dtype).
        diffusers/models/vae.py:127-144 (Encoder.forward) + :567 (quant_conv)."""
        ops.require_cuda(x, "AutoencoderKL.encode")
        boc = self._cfg["block_out_channels"]
        x = ops.conv3x3(x, self._conv_w("encoder.conv_in.weight"), bias=self._f("encoder.conv_in.bias"))
        for i in range(len(boc)):
            for j in range(self._cfg["layers_per_block"]):
                x = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", x)
            if i < len(boc) - 1:      # Downsample2D(padding=0): F.pad (0, 1, 0, 1) + valid stride-2 conv (resnet.py:183-188)
                p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                x = ops.conv3x3(x, self._conv_w(p + ".weight"), bias=self._f(p + ".bias"), stride=2, pad_mode=1)
        x = self._resnet("encoder.mid_block.resnets.0", x)
        x = self._attn("encoder.mid_block.attentions.0", x)
        x = self._resnet("encoder.mid_block.resnets.1", x)
        x = self._gn("encoder.conv_norm_out", x, True)
        x = ops.conv3x3(x, self._conv_w("encoder.conv_out.weight"), bias=self._f("encoder.conv_out.bias"))
        NB, H, W, c2 = x.shape
        return ops.gemm(x.view(-1, c2), self._w1x1("quant_conv.weight"), bias=self._f("quant_conv.bias")).view(NB, H, W, c2)

    @torch.no_grad()
    def encode(self, x, return_dict=True):
        """diffusers/models/vae.py:565-573: x (n, 3, H, W) -> .latent_dist (DiagonalGaussianDistribution over (n, 4, H/8, W/8)),
        moments in fp32."""
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        ops.require_cuda(x, "AutoencoderKL.encode")
        n, c, h, w = x.shape
        if h % 8 or w % 8:
            raise ValueError(f"AutoencoderKL.encode: image size {h}x{w} must be a multiple of 8")
        xx = ops.ncfhw_to_nfhwc(x.view(n, c, 1, h, w), self._compute_dtype).view(n, h, w, c)
        m = self.encode_nhwc(xx)
        moments = ops.nfhwc_to_ncfhw(m.view(n, 1, m.shape[1], m.shape[2], m.shape[3])).view(n, m.shape[3], m.shape[1], m.shape[2])
        posterior = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (posterior,)
        return AutoencoderKLOutput(latent_dist=posterior)

    # ------------------------------------------------------------------------------------------ decoder engine
    def _gn(self, p, x, silu):
        return ops.groupnorm(x, self._f(p + ".weight"), self._f(p + ".bias"), self._cfg["norm_num_groups"], 1e-6, silu=silu)

    def _resnet(self, p, x):
        NB, H, W, Cin = x.shape
        h = self._gn(p + ".norm1", x, True)
        h = ops.conv3x3(h, self._conv_w(p + ".conv1.weight"), bias=self._f(p + ".conv1.bias"))
        h = self._gn(p + ".norm2", h, True)
        res = x
        if self._has(p + ".conv_shortcut.weight"):
            res = ops.gemm(x.view(-1, Cin), self._w1x1(p + ".conv_shortcut.weight"), bias=self._f(p + ".conv_shortcut.bias")).view(NB, H, W, -1)
        return ops.conv3x3(h, self._conv_w(p + ".conv2.weight"), bias=self._f(p + ".conv2.bias"), residual=res)

    def _attn(self, p, x):
        NB, H, W, C = x.shape
        HW = H * W
        t = self._gn(p + ".group_norm", x, False).view(NB * HW, C)
        q = ops.gemm(t, self._w(p + ".query.weight"), bias=self._f(p + ".query.bias")).view(NB, HW, C)
        k = ops.gemm(t, self._w(p + ".key.weight"), bias=self._f(p + ".key.bias")).view(NB, HW, C)
        # V^T[n] = Wv @ t[n]^T (bias folded into the P V GEMM: softmax rows sum to 1, so P (V + 1 b^T) = P V + b^T)
        wv = self._w(p + ".value.weight")
        vt = ops.gemm(wv.unsqueeze(0).expand(NB, C, C), t.view(NB, HW, C))            # [NB, C, HW]
        # one head of width C = 512 (attention.py:331-379, fp32 softmax :366): TMEM cannot hold a 512-wide O accumulator next to S, so the
        # scores go through memory - but only for a few frames at a time (<= 256 MB of fp32 scores live, instead of [NB, HW, HW] at once:
        # 1.07 GB at 16 frames of 512x512, 10.9 GB at 32 frames of 768x768)
        o = torch.empty((NB, HW, C), dtype=x.dtype, device=x.device)
        fc = max(1, (1 << 26) // (HW * HW))
        for f0 in range(0, NB, fc):
            f1 = min(NB, f0 + fc)
            scores = ops.gemm(q[f0:f1], k[f0:f1], alpha=C ** -0.5, out_f32=True)       # [fc, HW, HW] fp32
            probs = ops.softmax_rows(scores, x.dtype)
            del scores
            ops.gemm(probs, vt[f0:f1], bias=self._f(p + ".value.bias"), out=o[f0:f1])  # [fc, HW, C]
            del probs
        out = ops.gemm(o.view(NB * HW, C), self._w(p + ".proj_attn.weight"), bias=self._f(p + ".proj_attn.bias"),
                       residual=x.view(NB * HW, C))
        return out.view(NB, H, W, C)

    def decode_nhwc(self, z):
        """z [N, h, w, latent] channels-last in the compute dtype -> [N, 8h, 8w, 3] (compute dtype; possibly a [..., :3] view
        of a 16-channel buffer, which ops.frames_finalize / ops.nfhwc_to_ncfhw read through their channel stride)."""
        ops.require_cuda(z, "AutoencoderKL.decode")
        NB, H, W, lc = z.shape
        boc = self._cfg["block_out_channels"]
        x = ops.gemm(z.reshape(-1, lc), self._w1x1("post_quant_conv.weight"), bias=self._f("post_quant_conv.bias")).view(NB, H, W, lc)
        x = ops.conv3x3(x, self._conv_w("decoder.conv_in.weight"), bias=self._f("decoder.conv_in.bias"))
        x = self._resnet("decoder.mid_block.resnets.0", x)
        x = self._attn("decoder.mid_block.attentions.0", x)
        x = self._resnet("decoder.mid_block.resnets.1", x)
        for i in range(len(boc)):
            for j in range(self._cfg["layers_per_block"] + 1):
                x = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", x)
            if i < len(boc) - 1:
                p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                x = ops.conv3x3(x, self._conv_w(p + ".weight"), bias=self._f(p + ".bias"), upsample=2, w_phases=self._conv_w_up2(p + ".weight"))
        x = self._gn("decoder.conv_norm_out", x, True)
        w_out, b_out, cout = self._conv_head("decoder.conv_out", x.shape[0] * x.shape[1] * x.shape[2])
        return ops.conv3x3(x, w_out, bias=b_out)[..., :cout]

    @torch.no_grad()
    def decode(self, z, return_dict=True):
        """diffusers/models/vae.py:600-610: z (n, 4, h, w) -> .sample (n, 3, 8h, 8w) fp32."""
        z = z.to(device=self.device, dtype=torch.float32).contiguous()
        n, c, h, w = z.shape
        zz = ops.ncfhw_to_nfhwc(z.view(n, c, 1, h, w), self._compute_dtype).view(n, h, w, c)
        y = self.decode_nhwc(zz)
        out = ops.nfhwc_to_ncfhw(y.view(n, 1, y.shape[1], y.shape[2], y.shape[3]))
        out = out.view(n, y.shape[3], y.shape[1], y.shape[2])
        if not return_dict:
            return (out,)
        return DecoderOutput(sample=out)

    def forward(self, sample, sample_posterior=False, return_dict=True, generator=None):
        """diffusers/models/vae.py:612-638."""
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        dec = self.decode(z).sample
        if not return_dict:
            return (dec,)
        return DecoderOutput(sample=dec)
