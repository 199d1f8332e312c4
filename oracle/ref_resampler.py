"""Oracle: functional fp32 restatement of the IP-Adapter image-prompt projectors.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``Resampler`` is the Perceiver resampler behind ``MyIPAdapterPlus``
(ip_adapter/my_ip_adapter.py:218-232 builds it with depth 4, 12 heads x 64, ``num_tokens`` queries); the UNet calls it on every
forward (animatediff/models/unet.py:592-594) although its input is constant over the DDIM loop.
State-dict keys are the reference module's: ``latents``, ``proj_in.*``, ``proj_out.*``, ``norm_out.*``,
``layers.{i}.0.{norm1,norm2,to_q,to_kv,to_out}.*`` (PerceiverAttention), ``layers.{i}.1.{0,1,3}.*`` (LayerNorm, Linear, GELU, Linear).
"""
import torch
import torch.nn.functional as F


def default_resampler_config(**over):
    """MyIPAdapterPlus.init_proj (ip_adapter/my_ip_adapter.py:221-231) for SD-1.5 + CLIP ViT-H/14 features."""
    cfg = dict(dim=768, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=1280, output_dim=768, ff_mult=4)
    cfg.update(over)
    return cfg


def resampler_param_shapes(cfg):
    d, inner = cfg["dim"], cfg["dim_head"] * cfg["heads"]
    s = {"latents": (1, cfg["num_queries"], d), "proj_in.weight": (d, cfg["embedding_dim"]), "proj_in.bias": (d,),
         "proj_out.weight": (cfg["output_dim"], d), "proj_out.bias": (cfg["output_dim"],),
         "norm_out.weight": (cfg["output_dim"],), "norm_out.bias": (cfg["output_dim"],)}
    for i in range(cfg["depth"]):
        p = f"layers.{i}"
        for n in ("norm1", "norm2"):
            s[f"{p}.0.{n}.weight"] = (d,); s[f"{p}.0.{n}.bias"] = (d,)
        s[f"{p}.0.to_q.weight"] = (inner, d); s[f"{p}.0.to_kv.weight"] = (2 * inner, d); s[f"{p}.0.to_out.weight"] = (d, inner)
        s[f"{p}.1.0.weight"] = (d,); s[f"{p}.1.0.bias"] = (d,)
        s[f"{p}.1.1.weight"] = (d * cfg["ff_mult"], d); s[f"{p}.1.3.weight"] = (d, d * cfg["ff_mult"])
    return s


def perceiver_attention(sd, p, cfg, x, latents):
    """ip_adapter/resampler.py:50-84: LN both inputs, q from the latents, k/v from [x ; latents], softmax in fp32 with the
    d^-1/4 scale applied to q and k separately (:74-76), heads merged back, bias-free output projection."""
    d = cfg["dim"]
    x = F.layer_norm(x, (d,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    latents = F.layer_norm(latents, (d,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])
    b, l, _ = latents.shape
    h = cfg["heads"]
    q = F.linear(latents, sd[p + ".to_q.weight"])
    k, v = F.linear(torch.cat((x, latents), dim=-2), sd[p + ".to_kv.weight"]).chunk(2, dim=-1)
    split = lambda t: t.view(b, t.shape[1], h, -1).transpose(1, 2)
    q, k, v = split(q), split(k), split(v)
    scale = cfg["dim_head"] ** -0.25
    w = torch.softmax(((q * scale) @ (k * scale).transpose(-2, -1)).float(), dim=-1)
    out = (w @ v).permute(0, 2, 1, 3).reshape(b, l, -1)
    return F.linear(out, sd[p + ".to_out.weight"])


def resampler_forward(sd, cfg, x):
    """ip_adapter/resampler.py:137-158 (apply_pos_emb False, no mean-pooled latents - the MyIPAdapterPlus construction)."""
    d = cfg["dim"]
    latents = sd["latents"].repeat(x.shape[0], 1, 1)
    x = F.linear(x, sd["proj_in.weight"], sd["proj_in.bias"])
    for i in range(cfg["depth"]):
        p = f"layers.{i}"
        latents = perceiver_attention(sd, p + ".0", cfg, x, latents) + latents
        hdn = F.layer_norm(latents, (d,), sd[p + ".1.0.weight"], sd[p + ".1.0.bias"])
        hdn = F.linear(F.gelu(F.linear(hdn, sd[p + ".1.1.weight"])), sd[p + ".1.3.weight"])
        latents = hdn + latents
    latents = F.linear(latents, sd["proj_out.weight"], sd["proj_out.bias"])
    return F.layer_norm(latents, (cfg["output_dim"],), sd["norm_out.weight"], sd["norm_out.bias"])


def image_proj_forward(sd, cross_attention_dim, tokens, image_embeds):
    """ImageProjModel (ip_adapter/my_ip_adapter.py:28-45): Linear -> reshape (b, T, D) -> LayerNorm."""
    y = F.linear(image_embeds, sd["proj.weight"], sd["proj.bias"]).reshape(-1, tokens, cross_attention_dim)
    return F.layer_norm(y, (cross_attention_dim,), sd["norm.weight"], sd["norm.bias"])
