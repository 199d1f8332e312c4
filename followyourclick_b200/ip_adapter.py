"""IP-Adapter surface: ``IPAttnProcessor`` (ip_adapter/attention_processor.py:80-183) and ``MyIPAdapter``
(ip_adapter/my_ip_adapter.py:47-134), engine underneath.

Inside ``UNet3DConditionModel`` the same math runs as part of the fused transformer block (unet.py::_transformer,
mirroring IPCrossAttention.forward, animatediff/models/attention.py:49-127).  ``IPAttnProcessor`` is the standalone,
processor-style entry for callers that drive an attention layer themselves: identical arithmetic
``softmax(q k_text^T s) v_text + scale * softmax(q k_ip^T s) v_ip`` followed by ``to_out``, on the same kernels.
"""
import torch
from torch import nn

from . import ops
from .unet import ImageProjModel


class IPAttnProcessor(nn.Module):
    def __init__(self, hidden_size, cross_attention_dim=None, scale=1.0, num_tokens=4):
        super().__init__()
        self.hidden_size, self.cross_attention_dim, self.scale, self.num_tokens = hidden_size, cross_attention_dim, scale, num_tokens
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)

    @torch.no_grad()
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        """attn: object with to_q / to_k / to_v / to_out[0] (nn.Linear-like: .weight, .bias) and .heads (and optional
        .scale = d^-1/2).  hidden_states (B, L, C) or (B, C, H, W); encoder_hidden_states (B, 77+T, Dc)."""
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is None on the whole inference path (SURVEY App. A.4)")
        if not hidden_states.is_cuda:
            raise RuntimeError("IPAttnProcessor runs only on CUDA (B200)")
        nd = hidden_states.dim()
        x = hidden_states
        if nd == 4:
            b, c, h, w = x.shape
            x = x.view(b, c, h * w).transpose(1, 2)
        dt = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.bfloat16
        x = x.to(dt).contiguous()
        B, Lq, C = x.shape
        heads = attn.heads
        d = C // heads
        sc = float(getattr(attn, "scale", d ** -0.5))
        w = lambda lin: lin.weight.detach().to(device=x.device, dtype=dt).contiguous()
        f = lambda t: t.detach().to(device=x.device, dtype=torch.float32).contiguous()
        q = ops.gemm(x.view(B * Lq, C), w(attn.to_q)).view(B, Lq, C)
        if encoder_hidden_states is None:
            ctx, ip = x, None
        else:
            e = encoder_hidden_states.to(device=x.device, dtype=dt).contiguous()
            end = e.shape[1] - self.num_tokens
            ctx, ip = e[:, :end].contiguous(), e[:, end:].contiguous()
        Bc, L, Dc = ctx.shape
        k = ops.gemm(ctx.view(Bc * L, Dc), w(attn.to_k)).view(Bc, L, C)
        v = ops.gemm(ctx.view(Bc * L, Dc), w(attn.to_v)).view(Bc, L, C)
        o = ops.attention(q, k, v, heads, sc)
        if ip is not None:
            T = ip.shape[1]
            ki = ops.gemm(ip.view(Bc * T, Dc), w(self.to_k_ip)).view(Bc, T, C)
            vi = ops.gemm(ip.view(Bc * T, Dc), w(self.to_v_ip)).view(Bc, T, C)
            ops.attention(q, ki, vi, heads, sc, out=o, out_alpha=float(self.scale), accumulate=True)
        out_lin = attn.to_out[0]
        y = ops.gemm(o.view(B * Lq, C), w(out_lin), bias=f(out_lin.bias) if out_lin.bias is not None else None).view(B, Lq, C)
        if nd == 4:
            y = y.transpose(1, 2).reshape(b, c, h, w)
        return y.to(hidden_states.dtype)


IPAttnProcessor2_0 = IPAttnProcessor


class MyIPAdapter:
    """ip_adapter/my_ip_adapter.py:47-134.  The CLIP vision tower is a per-clip one-off outside the hot path: pass any
    ``image_encoder`` callable returning ``.image_embeds`` (transformers.CLIPVisionModelWithProjection or a stub)."""

    def __init__(self, unet, image_encoder_path=None, ip_ckpt=None, device="cuda", num_tokens=4, image_encoder=None,
                 clip_embeddings_dim=None):
        self.device, self.image_encoder_path, self.ip_ckpt, self.num_tokens, self.unet = device, image_encoder_path, ip_ckpt, num_tokens, unet
        if image_encoder is None and image_encoder_path:
            from transformers import CLIPVisionModelWithProjection
            image_encoder = CLIPVisionModelWithProjection.from_pretrained(image_encoder_path).to(device)
        self.image_encoder = image_encoder
        self._clip_dim = clip_embeddings_dim or getattr(getattr(image_encoder, "config", None), "projection_dim", 1024)
        self.clip_image_processor = None
        self.image_proj_model = self.init_proj()

    def init_proj(self):
        return ImageProjModel(cross_attention_dim=self.unet.config.cross_attention_dim, clip_embeddings_dim=self._clip_dim,
                              clip_extra_context_tokens=self.num_tokens).to(self.device)

    @torch.no_grad()
    def get_image_clip_feat(self, input_image=None):
        if not torch.is_tensor(input_image):
            if self.clip_image_processor is None:
                from transformers import CLIPImageProcessor
                self.clip_image_processor = CLIPImageProcessor()
            imgs = input_image if isinstance(input_image, list) else [input_image]
            input_image = self.clip_image_processor(images=imgs, return_tensors="pt").pixel_values
        emb = self.image_encoder(input_image.to(self.device)).image_embeds
        return emb, torch.zeros_like(emb)
