"""Shared test configurations ("mini" = the reference architecture at reduced width/depth).

The mini UNet keeps the head dims of the real model (d = 40 / 80 / 160 / 160) by using 4 heads on
(160, 320, 640, 640) channels, so the same kernel instantiations run in the tests as at full size.
"""
import torch

from oracle.ref_ddim import default_scheduler_config
from oracle.ref_unet import default_unet_config
from oracle.ref_vae import default_vae_config
from oracle.ref_resampler import default_resampler_config

CLIP_DIM = 1024
MINI_UNET_VARIANTS = ("base", "ip", "cam")
MINI_VAE = default_vae_config(block_out_channels=(32, 64, 128, 128), layers_per_block=1)
SCHED_V = default_scheduler_config()
# Perceiver resampler at reduced size: keeps dim_head 64 (the real head size) and an odd image-token count like CLIP's 257
MINI_RESAMPLER = default_resampler_config(dim=128, depth=2, dim_head=64, heads=3, num_queries=8, embedding_dim=96, output_dim=768)
RESAMPLER_TOKENS = 33
SCHED_EPS = default_scheduler_config(prediction_type="epsilon", rescale_betas_zero_snr=False)

_MM = dict(num_attention_heads=4, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
           temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1,
           zero_initialize=True)


def mini_unet_ref_kwargs(variant):
    """Constructor kwargs accepted both by the reference UNet3DConditionModel (unet.py:43-103) and by ours."""
    kw = dict(sample_size=16, in_channels=4, out_channels=4, block_out_channels=(160, 320, 640, 640),
              layers_per_block=1, attention_head_dim=4, cross_attention_dim=768, norm_num_groups=32,
              use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), unet_use_cross_frame_attention=False,
              unet_use_temporal_attention=False, motion_module_type="Vanilla", motion_module_kwargs=dict(_MM))
    if variant == "base":          # configs/inference/inference_img_embed_mask_condition_zero_snr_.yaml
        kw.update(use_fps_condition=True, use_first_frame_mask_condition_concat=True)
    elif variant == "ip":          # cfg3: base + IP-Adapter cross attention
        kw.update(use_fps_condition=True, use_first_frame_mask_condition_concat=True,
                  use_ip_cross_attention=True, scale=0.25, num_tokens=4)
    elif variant == "cam":         # cfg5: training_magic_..._lora_32f_from_244k.yaml:7-38
        mm = dict(_MM, temporal_position_encoding_max_len=32, add_temporal_lora=True, rank=4)
        kw.update(use_ip_cross_attention=True, image_condition_dim=1024, scale=0.2, num_tokens=4,
                  use_camera_motion_condition=True, motion_module_kwargs=mm, layers_per_block=2)
    else:
        raise KeyError(variant)
    return kw


def mini_unet_oracle_cfg(variant):
    kw = mini_unet_ref_kwargs(variant)
    mm = {k: v for k, v in kw["motion_module_kwargs"].items() if k != "zero_initialize"}
    return default_unet_config(
        block_out_channels=kw["block_out_channels"], layers_per_block=kw["layers_per_block"],
        attention_head_dim=kw["attention_head_dim"], motion_module_kwargs=mm,
        use_first_frame_mask_condition_concat=kw.get("use_first_frame_mask_condition_concat", False),
        use_fps_condition=kw.get("use_fps_condition", False),
        use_ip_cross_attention=kw.get("use_ip_cross_attention", False),
        scale=kw.get("scale", 1.0), num_tokens=kw.get("num_tokens", 4),
        use_camera_motion_condition=kw.get("use_camera_motion_condition", False))


def unet_inputs(variant, b=2, f=4, h=16, w=16, seed=11):
    g = torch.Generator().manual_seed(seed)
    cin = 9 if variant in ("base", "ip") else 4
    d = dict(sample=torch.randn(b, cin, f, h, w, generator=g), timestep=torch.tensor(501),
             ctx=torch.randn(b, 77, 768, generator=g))
    if variant in ("base", "ip"):
        d["fps"] = torch.tensor([2, 2])
        d["flow"] = torch.tensor([4, 4])
    if variant in ("ip", "cam"):
        d["clip"] = torch.randn(b, CLIP_DIM, generator=g)
    if variant == "cam":
        d["camera"] = torch.tensor([3, 3])
    return d


# stock 2-D UNet (diffusers UNet2DConditionModel) at the mini width: the T2I first-frame generator
MINI_UNET2D = dict(sample_size=16, in_channels=4, out_channels=4, block_out_channels=(160, 320, 640, 640), layers_per_block=1,
                   attention_head_dim=4, cross_attention_dim=768, norm_num_groups=32)


def mini_unet2d_oracle_cfg():
    return default_unet_config(block_out_channels=MINI_UNET2D["block_out_channels"], layers_per_block=1, attention_head_dim=4,
                               use_motion_module=False, use_first_frame_mask_condition_concat=False, use_fps_condition=False)


def pipeline_variant_inputs(variant):
    """Pipeline-level plumbing cases beyond the shipped YAML: returns (clip inputs, reference-pipeline kwargs, oracle-denoise kwargs,
    scheduler config, steps, guidance)."""
    from followyourclick_b200.synth import synth_clip_inputs
    ci = synth_clip_inputs(1, 4, 8, 8, clip_dim=CLIP_DIM)
    ci["uncond_image_clip_feat"] = torch.randn(1, CLIP_DIM, generator=torch.Generator().manual_seed(77)) * 0.1   # a non-zero uncond feature: ordering bugs show
    if variant == "ip":        # cfg3: shipped YAML + IP-Adapter image condition
        kw = dict(use_first_frame_mask_condition_concat=True, first_image_latents=ci["first_image_latents"], use_fps_condition=True,
                  fps_tensor=torch.tensor([2]), flow_control=torch.tensor([4]), first_images_mask=ci["first_images_mask"])
        okw = dict(first_image_latents=ci["first_image_latents"], first_images_mask=ci["first_images_mask"], fps_tensor=torch.tensor([2]),
                   flow_control=torch.tensor([4]))
        return ci, kw, okw, SCHED_V, 2, 8.0
    if variant == "cam":       # cfg5: camera-LoRA model, 4-channel input, epsilon prediction
        kw = dict(use_camera_motion_condition=True, camera_movement_type=torch.tensor([3]))
        okw = dict(camera_movement_type=torch.tensor([3]))
        return ci, kw, okw, SCHED_EPS, 2, 7.5
    raise KeyError(variant)


# IPAttnProcessor cases (tests/golden/ip_attn_processor.npz holds only the reference processor's OUTPUTS; weights and inputs are
# regenerated from the case name, so the fixture stays small): name -> (C, heads, T image tokens, 4-D input spatial shape | None, residual)
IP_ATTN_CASES = {"tok_t4": (160, 4, 4, None, False), "img_t16": (320, 4, 16, (8, 8), False), "res_t4": (160, 2, 4, None, True)}
IP_ATTN_SCALE = 0.7


def ip_attn_case(name, xd=768, B=2, L=64):
    import zlib
    C, heads, T, shape4d, residual = IP_ATTN_CASES[name]
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    rn = lambda *s_, sc=1.0: torch.randn(*s_, generator=g) * sc
    w = dict(to_q=rn(C, C, sc=C ** -0.5), to_k=rn(C, xd, sc=xd ** -0.5), to_v=rn(C, xd, sc=xd ** -0.5), to_out_w=rn(C, C, sc=C ** -0.5),
             to_out_b=rn(C, sc=0.05), to_k_ip=rn(C, xd, sc=xd ** -0.5), to_v_ip=rn(C, xd, sc=xd ** -0.5))
    x = rn(B, C, *shape4d) if shape4d else rn(B, L, C)
    ctx = rn(B, 77 + T, xd)
    return dict(C=C, heads=heads, T=T, shape4d=shape4d, residual=residual, xd=xd, w=w, x=x, ctx=ctx)


class DuckAttention(torch.nn.Module):
    """What an IP-Adapter attention processor is handed: the ``Attention`` module of modern pip diffusers (>= 0.19; NOT the vendored
    0.11.1 CrossAttention, and absent from /root/reference), restated from its published helpers in plain torch.  Used to drive the
    UNMODIFIED reference processor when the fixture is generated and to drive the engine's processor in the tests."""

    def __init__(self, case):
        super().__init__()
        C, xd, w = case["C"], case["xd"], case["w"]
        self.heads, self.scale = case["heads"], (C // case["heads"]) ** -0.5
        self.to_q, self.to_k, self.to_v = torch.nn.Linear(C, C, bias=False), torch.nn.Linear(xd, C, bias=False), torch.nn.Linear(xd, C, bias=False)
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(C, C), torch.nn.Dropout(0.0)])
        self.spatial_norm = self.group_norm = None
        self.norm_cross, self.residual_connection, self.rescale_output_factor = False, case["residual"], 1.0
        with torch.no_grad():
            self.to_q.weight.copy_(w["to_q"]); self.to_k.weight.copy_(w["to_k"]); self.to_v.weight.copy_(w["to_v"])
            self.to_out[0].weight.copy_(w["to_out_w"]); self.to_out[0].bias.copy_(w["to_out_b"])

    def prepare_attention_mask(self, mask, target_length, batch_size):
        return mask

    def head_to_batch_dim(self, t):
        b, L, c = t.shape
        return t.reshape(b, L, self.heads, c // self.heads).permute(0, 2, 1, 3).reshape(b * self.heads, L, c // self.heads)

    def batch_to_head_dim(self, t):
        bh, L, d = t.shape
        return t.reshape(bh // self.heads, self.heads, L, d).permute(0, 2, 1, 3).reshape(bh // self.heads, L, d * self.heads)

    def get_attention_scores(self, q, k, mask=None):
        assert mask is None
        return torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], device=q.device, dtype=q.dtype), q, k.transpose(-1, -2),
                             beta=0, alpha=self.scale).softmax(dim=-1)
