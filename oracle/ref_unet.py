"""Oracle: functional fp32 restatement of ``UNet3DConditionModel.forward``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Operates on a flat state dict with the
reference key names and a plain config dict; every function cites the
reference file:line (relative to /root/reference) that it restates.  Tensors keep the
reference layout (b, c, f, h, w) so intermediate values are directly comparable with
hooks on the reference model.
"""
import math

import torch
import torch.nn.functional as F


def default_unet_config(**over):
    """SD-1.5 hyper-parameters + configs/inference/inference_img_embed_mask_condition_zero_snr_.yaml:1-26."""
    cfg = dict(
        in_channels=4, out_channels=4,
        block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
        attention_head_dim=8,            # number of heads in this diffusers vintage (unet.py:212)
        cross_attention_dim=768, norm_num_groups=32, norm_eps=1e-5,
        flip_sin_to_cos=True, freq_shift=0,
        use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8),
        motion_module_mid_block=False, motion_module_decoder_only=False,
        motion_module_kwargs=dict(
            num_attention_heads=8, num_transformer_block=1,
            attention_block_types=("Temporal_Self", "Temporal_Self"),
            temporal_position_encoding=True, temporal_position_encoding_max_len=24,
            temporal_attention_dim_div=1, add_temporal_lora=False, rank=4),
        use_first_frame_condition_concat=False, use_first_frame_mask_condition_concat=True,
        use_ip_cross_attention=False, scale=1.0, num_tokens=4,
        use_fps_condition=True, use_camera_motion_condition=False,
        use_inflated_groupnorm=False,
    )
    mm = dict(cfg["motion_module_kwargs"])
    mm.update(over.pop("motion_module_kwargs", {}))
    cfg.update(over)
    cfg["motion_module_kwargs"] = mm
    return cfg


def _heads(cfg, i):
    h = cfg["attention_head_dim"]
    return h[i] if isinstance(h, (tuple, list)) else h


# ----------------------------------------------------------------------------- primitives

def timestep_sinusoid(t, dim, flip_sin_to_cos, freq_shift):
    """diffusers/models/embeddings.py:21-61 (scale=1, max_period=10000)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32) / (half - freq_shift)
    ang = t[:, None].float() * torch.exp(exponent)[None, :].to(t.device)      # table evaluated on the host like the reference
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def timestep_mlp(sd, p, x):
    """diffusers/models/embeddings.py:64-88: Linear -> SiLU -> Linear."""
    x = F.linear(x, sd[p + ".linear_1.weight"], sd[p + ".linear_1.bias"])
    return F.linear(F.silu(x), sd[p + ".linear_2.weight"], sd[p + ".linear_2.bias"])


def conv2d_per_frame(sd, p, x, stride=1, padding=1):
    """animatediff/models/resnet.py:19-27 InflatedConv3d: conv2d applied to every frame."""
    b, c, f, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    y = F.conv2d(y, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)
    return y.reshape(b, f, *y.shape[1:]).permute(0, 2, 1, 3, 4)


def group_norm_5d(sd, p, x, groups, eps, per_frame):
    """Cross-frame statistics (nn.GroupNorm on the 5-D tensor, resnet.py:240,263) or
    per-frame statistics (InflatedGroupNorm resnet.py:9-17; attention.py:269; motion_module.py:188)."""
    if not per_frame:
        return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)
    b, c, f, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    y = F.group_norm(y, groups, sd[p + ".weight"], sd[p + ".bias"], eps)
    return y.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def layer_norm(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


MHA_BATCH_CHUNK = None      # bench.py's CPU leg sets this: evaluate the batch entries in chunks (rows are independent, so the result is
                            # the same arithmetic) - the reference materialises (B*heads, Lq, Lk) scores, 17 GB at the cfg2 level-0 shape


def mha(q, k, v, heads, scale=None):
    """diffusers/models/attention.py:649-678 (== mm_attn_cross.py:148-177): softmax(q k^T * scale) v
    per head, heads packed along the channel axis; scale defaults to d^-1/2 (attention.py:544)."""
    if MHA_BATCH_CHUNK and q.shape[0] > MHA_BATCH_CHUNK:
        n = MHA_BATCH_CHUNK
        return torch.cat([_mha(q[i:i + n], k[i:i + n], v[i:i + n], heads, scale) for i in range(0, q.shape[0], n)])
    return _mha(q, k, v, heads, scale)


def _mha(q, k, v, heads, scale=None):
    B, Lq, C = q.shape
    d = C // heads
    scale = d ** -0.5 if scale is None else scale
    qh = q.reshape(B, Lq, heads, d).permute(0, 2, 1, 3)
    kh = k.reshape(B, k.shape[1], heads, d).permute(0, 2, 1, 3)
    vh = v.reshape(B, v.shape[1], heads, d).permute(0, 2, 1, 3)
    s = torch.matmul(qh, kh.transpose(-1, -2)) * scale
    o = torch.matmul(s.softmax(dim=-1), vh)
    return o.permute(0, 2, 1, 3).reshape(B, Lq, C)


def feed_forward(sd, p, x):
    """diffusers/models/attention.py:733-775,800-821: GEGLU (exact-erf gelu) then Linear."""
    h = F.linear(x, sd[p + ".net.0.proj.weight"], sd[p + ".net.0.proj.bias"])
    a, gate = h.chunk(2, dim=-1)
    return F.linear(a * F.gelu(gate), sd[p + ".net.2.weight"], sd[p + ".net.2.bias"])


# ----------------------------------------------------------------------------- blocks

def resnet_block_3d(sd, p, x, emb, cfg):
    """animatediff/models/resnet.py:296-342."""
    g, eps, pf = cfg["norm_num_groups"], cfg["norm_eps"], cfg["use_inflated_groupnorm"]
    h = F.silu(group_norm_5d(sd, p + ".norm1", x, g, eps, pf))
    h = conv2d_per_frame(sd, p + ".conv1", h)
    t = F.linear(F.silu(emb), sd[p + ".time_emb_proj.weight"], sd[p + ".time_emb_proj.bias"])
    h = h + t[:, :, None, None, None]
    h = F.silu(group_norm_5d(sd, p + ".norm2", h, g, eps, pf))
    h = conv2d_per_frame(sd, p + ".conv2", h)
    if (p + ".conv_shortcut.weight") in sd:
        x = conv2d_per_frame(sd, p + ".conv_shortcut", x, padding=0)
    return x + h


def cross_attention(sd, p, x, ctx, heads, cfg):
    """CrossAttention.forward diffusers/models/attention.py:592-647, or IPCrossAttention.forward
    animatediff/models/attention.py:49-127 when the IP projections exist."""
    q = F.linear(x, sd[p + ".to_q.weight"])
    if ctx is None:
        ctx = x
    ip, qk_scale = None, None
    if (p + ".to_k_ip.weight") in sd:
        T = cfg["num_tokens"]
        ctx, ip = ctx[:, :-T], ctx[:, -T:]
        # REFERENCE QUIRK (animatediff/models/attention.py:43): IPCrossAttention.__init__ overwrites
        # CrossAttention.scale (= d^-1/2, diffusers attention.py:544) with the IP-adapter scale, so the
        # non-xformers `_attention` (attention.py:98,115 -> baddbmm alpha=self.scale) multiplies the logits of
        # BOTH softmaxes by the IP scale instead of d^-1/2.  The xformers path ignores self.scale (uses d^-1/2).
        # The oracle follows the executed CPU path unless cfg["xformers_semantics"] is set.
        if not cfg.get("xformers_semantics", False):
            qk_scale = cfg["scale"]
    o = mha(q, F.linear(ctx, sd[p + ".to_k.weight"]), F.linear(ctx, sd[p + ".to_v.weight"]), heads, qk_scale)
    if ip is not None:
        o_ip = mha(q, F.linear(ip, sd[p + ".to_k_ip.weight"]), F.linear(ip, sd[p + ".to_v_ip.weight"]), heads, qk_scale)
        o = o + cfg["scale"] * o_ip
    return F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def transformer_3d(sd, p, x, ctx, heads, cfg):
    """animatediff/models/attention.py:217-308 + BasicTransformerBlock.forward :489-564."""
    b, c, f, h, w = x.shape
    res = x
    y = group_norm_5d(sd, p + ".norm", x, cfg["norm_num_groups"], 1e-6, per_frame=True)
    y = conv2d_per_frame(sd, p + ".proj_in", y, padding=0)
    tok = y.permute(0, 2, 3, 4, 1).reshape(b * f, h * w, c)
    ctx_f = ctx.repeat_interleave(f, dim=0)            # 'b n c -> (b f) n c'  (:264)
    i = 0
    while (p + f".transformer_blocks.{i}.norm1.weight") in sd:
        q = p + f".transformer_blocks.{i}"
        tok = tok + cross_attention(sd, q + ".attn1", layer_norm(sd, q + ".norm1", tok), None, heads, cfg)
        tok = tok + cross_attention(sd, q + ".attn2", layer_norm(sd, q + ".norm2", tok), ctx_f, heads, cfg)
        tok = tok + feed_forward(sd, q + ".ff", layer_norm(sd, q + ".norm3", tok))
        i += 1
    y = tok.reshape(b, f, h, w, c).permute(0, 4, 1, 2, 3)
    y = conv2d_per_frame(sd, p + ".proj_out", y, padding=0)
    return y + res


def _lora(sd, p, name, x):
    """motion_module.py:306-326,389-456 runtime LoRA (scale = 1.0)."""
    k = p + f".{name}_lora.down.weight"
    if k not in sd:
        return 0.0
    return F.linear(F.linear(x, sd[k]), sd[p + f".{name}_lora.up.weight"])


def temporal_attention(sd, p, xn, f, heads):
    """VersatileAttention.forward animatediff/models/motion_module.py:371-464 (Temporal_Self)."""
    BF, D, C = xn.shape
    b = BF // f
    h = xn.reshape(b, f, D, C).permute(0, 2, 1, 3).reshape(b * D, f, C)       # (b f) d c -> (b d) f c
    if (p + ".pos_encoder.pe") in sd:
        h = h + sd[p + ".pos_encoder.pe"][:, :f]
    q = F.linear(h, sd[p + ".to_q.weight"]) + _lora(sd, p, "to_q", h)
    k = F.linear(h, sd[p + ".to_k.weight"]) + _lora(sd, p, "to_k", h)
    v = F.linear(h, sd[p + ".to_v.weight"]) + _lora(sd, p, "to_v", h)
    o = mha(q, k, v, heads)
    o = F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"]) + _lora(sd, p, "to_out", o)
    return o.reshape(b, D, f, C).permute(0, 2, 1, 3).reshape(BF, D, C)


def motion_module(sd, p, x, cfg):
    """VanillaTemporalModule -> TemporalTransformer3DModel.forward motion_module.py:157-208,
    TemporalTransformerBlock.forward :270-283."""
    p = p + ".temporal_transformer"
    mm = cfg["motion_module_kwargs"]
    b, c, f, h, w = x.shape
    res = x
    y = group_norm_5d(sd, p + ".norm", x, 32, 1e-6, per_frame=True)
    tok = y.permute(0, 2, 3, 4, 1).reshape(b * f, h * w, c)
    tok = F.linear(tok, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    for i in range(mm["num_transformer_block"]):
        q = p + f".transformer_blocks.{i}"
        for j in range(len(mm["attention_block_types"])):
            tok = tok + temporal_attention(sd, q + f".attention_blocks.{j}",
                                           layer_norm(sd, q + f".norms.{j}", tok), f, mm["num_attention_heads"])
        tok = tok + feed_forward(sd, q + ".ff", layer_norm(sd, q + ".ff_norm", tok))
    tok = F.linear(tok, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return tok.reshape(b, f, h, w, c).permute(0, 4, 1, 2, 3) + res


def upsample_3d(sd, p, x):
    """animatediff/models/resnet.py:137-170: nearest (1,2,2) then 3x3 conv."""
    x = x.repeat_interleave(2, dim=3).repeat_interleave(2, dim=4)
    return conv2d_per_frame(sd, p + ".conv", x)


# ----------------------------------------------------------------------------- model

def image_proj(sd, clip_feat, cfg):
    """ip_adapter/my_ip_adapter.py:28-45 ImageProjModel (called each forward at unet.py:592-594)."""
    T, D = cfg["num_tokens"], cfg["cross_attention_dim"]
    y = F.linear(clip_feat, sd["image_proj_model.proj.weight"], sd["image_proj_model.proj.bias"]).reshape(-1, T, D)
    return F.layer_norm(y, (D,), sd["image_proj_model.norm.weight"], sd["image_proj_model.norm.bias"], 1e-5)


def unet3d_forward(sd, cfg, sample, timestep, encoder_hidden_states, fps_tensor=None, flow_control=None,
                   reference_images_clip_feat=None, camera_movement_type_tensor=None,
                   use_first_frame_condition_concat=False, reference_images_latent=None, taps=None):
    """animatediff/models/unet.py:422-672.  Returns the (b, out, f, h, w) prediction.
    ``taps`` (optional dict) receives named intermediate activations for layer-wise debugging."""
    sd = {k: v.float() for k, v in sd.items()}
    sample = sample.float()
    B = sample.shape[0]
    boc = cfg["block_out_channels"]
    n_lvl = len(boc)
    mmk = cfg["motion_module_kwargs"]

    def as_vec(v):
        v = torch.as_tensor(v).to(sample.device)      # device-agnostic: the -m gpu full-size parity tests run this oracle on cuda (fp32, TF32 off)
        return (v[None] if v.dim() == 0 else v).expand(B)

    def sinus(v):
        return timestep_sinusoid(as_vec(v), boc[0], cfg["flip_sin_to_cos"], cfg["freq_shift"])

    emb = timestep_mlp(sd, "time_embedding", sinus(timestep))                    # unet.py:522-533
    if cfg["use_camera_motion_condition"] and camera_movement_type_tensor is not None:
        emb = emb + timestep_mlp(sd, "camera_motion_embedding", sinus(camera_movement_type_tensor))  # :537-542
    if cfg["use_fps_condition"] and fps_tensor is not None:
        emb = emb + timestep_mlp(sd, "fps_embedding", sinus(fps_tensor))         # :545-551
        emb = emb + timestep_mlp(sd, "motion_embedding", sinus(flow_control))    # :554-558

    if use_first_frame_condition_concat and reference_images_latent is not None:     # :578-583
        first = reference_images_latent.float().unsqueeze(2).repeat(1, 1, sample.shape[2], 1, 1)
        sample = torch.cat((sample, first), dim=1)
    x = conv2d_per_frame(sd, "conv_in", sample)                                  # :586
    if use_first_frame_condition_concat:
        x = x / 2                                                                # :589-590
    ctx = encoder_hidden_states.float()
    if cfg["use_ip_cross_attention"] and reference_images_clip_feat is not None:
        ctx = torch.cat([ctx, image_proj(sd, reference_images_clip_feat.float(), cfg)], dim=1)   # :592-594
    if taps is not None:
        taps["conv_in"] = x

    def maybe_motion(p, x, res_index, decoder):
        on = cfg["use_motion_module"] and (2 ** res_index) in cfg["motion_module_resolutions"]
        if not decoder and cfg["motion_module_decoder_only"]:
            on = False
        return motion_module(sd, p, x, cfg) if on else x

    skips = [x]
    for i in range(n_lvl):                                                       # down, :601-626
        p = f"down_blocks.{i}"
        has_attn = i < n_lvl - 1
        for j in range(cfg["layers_per_block"]):
            x = resnet_block_3d(sd, f"{p}.resnets.{j}", x, emb, cfg)
            if has_attn:
                x = transformer_3d(sd, f"{p}.attentions.{j}", x, ctx, _heads(cfg, i), cfg)
            x = maybe_motion(f"{p}.motion_modules.{j}", x, i, decoder=False)
            skips.append(x)
        if i < n_lvl - 1:
            x = conv2d_per_frame(sd, f"{p}.downsamplers.0.conv", x, stride=2)    # resnet.py:184
            skips.append(x)
        if taps is not None:
            taps[f"down{i}"] = x

    x = resnet_block_3d(sd, "mid_block.resnets.0", x, emb, cfg)                  # unet_blocks.py:342-360
    x = transformer_3d(sd, "mid_block.attentions.0", x, ctx, _heads(cfg, n_lvl - 1), cfg)
    if cfg["use_motion_module"] and cfg["motion_module_mid_block"]:
        x = motion_module(sd, "mid_block.motion_modules.0", x, cfg)
    x = resnet_block_3d(sd, "mid_block.resnets.1", x, emb, cfg)
    if taps is not None:
        taps["mid"] = x

    for i in range(n_lvl):                                                       # up, :636-660
        p = f"up_blocks.{i}"
        has_attn = i > 0
        lvl = n_lvl - 1 - i
        for j in range(cfg["layers_per_block"] + 1):
            x = torch.cat([x, skips.pop()], dim=1)                               # unet_blocks.py:763,885
            x = resnet_block_3d(sd, f"{p}.resnets.{j}", x, emb, cfg)
            if has_attn:
                x = transformer_3d(sd, f"{p}.attentions.{j}", x, ctx, _heads(cfg, lvl), cfg)
            x = maybe_motion(f"{p}.motion_modules.{j}", x, lvl, decoder=True)
        if i < n_lvl - 1:
            x = upsample_3d(sd, f"{p}.upsamplers.0", x)
        if taps is not None:
            taps[f"up{i}"] = x

    x = F.silu(group_norm_5d(sd, "conv_norm_out", x, cfg["norm_num_groups"], cfg["norm_eps"],
                             cfg["use_inflated_groupnorm"]))                     # :665-666
    return conv2d_per_frame(sd, "conv_out", x)                                   # :667
