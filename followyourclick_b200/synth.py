"""Deterministic synthetic weights and inputs.

There are no pretrained checkpoints offline, so every parity test and the benchmark run on
procedurally generated weights.  Values depend only on (key name, shape, seed), never on module
construction order, so the reference model (in the build container), the CPU oracle and the CUDA
engine can all be loaded with the *same* state dict without shipping it.

Scaling keeps activations O(1) through the network with attention logits of unit scale (std ~1: non-uniform
softmaxes, so key-ordering / masking bugs are visible, yet well conditioned - with logit std ~4 the random network
amplifies a 1e-4 input perturbation 12x and bf16 rounding to ~9 % at the output, measured in tests/diag_bf16.py):
  * >=2-D ``*.weight``: N(0, 1/fan_in)
  * 1-D ``*.weight`` (norm gains): 1 + 0.1 N(0,1);  ``*.bias``: 0.05 N(0,1)
  * ``*.pos_encoder.pe`` buffers are left untouched (they are a formula, motion_module.py:295-299)
  * zero-initialised reference tensors (motion proj_out, fps/motion embedding linear_2) are drawn
    like every other tensor - otherwise those branches are no-ops and parity is vacuous.
"""
import math
import zlib

import torch

_QK_GAIN = 1.0


def synth_tensor(key, shape, seed=0):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(shape)
    x = torch.randn(shape, generator=g, dtype=torch.float32)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "bias":
        return 0.05 * x
    if len(shape) == 1:
        return 1.0 + 0.1 * x
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    gain = 1.0
    parent = key.rsplit(".", 2)[-2] if key.count(".") >= 1 else ""
    if parent in ("to_q", "to_k", "query", "key"):
        gain = _QK_GAIN
    if parent in ("down", "up"):          # LoRA factors: keep the low-rank delta comparable to W
        gain = 0.5 * math.sqrt(shape[1]) if parent == "up" else 1.0
    return x * (gain / math.sqrt(fan_in))


def synth_state_dict(shapes, seed=0, skip=(".pos_encoder.pe",)):
    """shapes: mapping key -> shape.  Returns {key: fp32 CPU tensor} for all keys not in ``skip``."""
    out = {}
    for k, shp in shapes.items():
        if any(k.endswith(s) for s in skip):
            continue
        out[k] = synth_tensor(k, shp, seed)
    return out


def load_synth_(module, seed=0):
    """Overwrite ``module``'s parameters/buffers in place with synthetic values (keeps pe buffers)."""
    sd = module.state_dict()
    new = synth_state_dict({k: v.shape for k, v in sd.items()}, seed)
    missing, unexpected = module.load_state_dict(new, strict=False)
    assert not unexpected, unexpected
    return module


def synth_clip_inputs(b, f, h, w, seed=1234, ctx_len=77, ctx_dim=768, clip_dim=None, dtype=torch.float32):
    """Synthetic per-clip inputs of SURVEY 8d: latents, first-frame latents, rectangle mask, text embeddings."""
    def rn(shape, s):
        return torch.randn(shape, generator=torch.Generator().manual_seed(s), dtype=torch.float32).to(dtype)
    out = dict(
        latents=rn((b, 4, f, h, w), seed),
        first_image_latents=rn((b, 4, h, w), seed + 1),
        text_embeddings=rn((2 * b, ctx_len, ctx_dim), seed + 2),
    )
    mask = torch.zeros(b, 1, 1, h, w, dtype=dtype)
    mask[..., h // 4: 3 * h // 4, w // 4: 3 * w // 4] = 1
    out["first_images_mask"] = mask
    if clip_dim is not None:
        out["image_clip_feat"] = rn((b, clip_dim), seed + 3)
        out["uncond_image_clip_feat"] = torch.zeros(b, clip_dim, dtype=dtype)
    return out


def synth_on_device_(module, seed=0):
    """Fast on-device variant for the benchmark (values differ from the CPU generator; same distributions).  Used where
    only timing matters: drawing 1.28 B normals on the host takes tens of seconds."""
    dev = next(module.parameters()).device
    g = torch.Generator(device=dev).manual_seed(seed)
    for k, p in module.state_dict(keep_vars=True).items():
        if k.endswith(".pe"):
            continue
        x = torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32)
        if k.endswith(".bias"):
            p.data = 0.05 * x
        elif p.dim() == 1:
            p.data = 1.0 + 0.1 * x
        else:
            p.data = x * (p[0].numel() ** -0.5)
    if hasattr(module, "_invalidate"):
        module._invalidate()
    return module
