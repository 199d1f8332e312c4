"""Oracle: fp32 restatement of the ``AnimationPipeline.__call__`` denoise loop + decode.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference:
animatediff/pipelines/pipeline_animation.py:620-635 (timesteps, mask broadcast), :686-773 (loop),
:779 + :400-413 (decode_latents).  Prompt/image encoders are outside the hot path: their outputs
(text_embeddings with batch order [uncond, cond]; CLIP image features) are inputs here.
"""
import torch

from .ref_ddim import DDIMOracle, cfg_combine
from .ref_unet import unet3d_forward
from .ref_vae import decode_latents


def build_unet_input(latents, first_image_latents, first_images_mask, use_mask_concat):
    """pipeline_animation.py:625-635,693-711: [latents(4) | mask(1) | first-frame block(4)], then CFG x2."""
    if not use_mask_concat:
        return torch.cat([latents] * 2)
    b, c, f, h, w = latents.shape
    block = torch.zeros_like(latents)
    block[:, :, 0] = first_image_latents
    if first_images_mask is not None:
        mask = torch.clamp(first_images_mask[:, :, 0:1].repeat(1, 1, f, 1, 1), 0, 1)
    else:
        mask = torch.zeros_like(latents)[:, :1]
        mask[:, :, 0] = 1
    x9 = torch.cat((latents, mask, block), dim=1)
    return torch.cat([x9] * 2)


def denoise(unet_sd, unet_cfg, sched_cfg, latents, text_embeddings, num_inference_steps, guidance_scale,
            first_image_latents=None, first_images_mask=None, fps_tensor=None, flow_control=None,
            image_clip_feat=None, uncond_image_clip_feat=None, camera_movement_type=None, trace=None, video_scale=0):
    """The hot loop: returns final latents (b, 4, f, h, w).  guidance_scale must be > 1 (CFG on).
    video_scale > 0: the per-frame guidance branch of pipeline_animation.py:738-761."""
    sched = DDIMOracle(sched_cfg)
    latents = latents.float().clone()
    use_concat = unet_cfg["use_first_frame_mask_condition_concat"]
    dup = lambda v: None if v is None else torch.cat([torch.as_tensor(v).reshape(-1)] * 2)
    clip = None
    if unet_cfg["use_ip_cross_attention"]:
        clip = torch.cat([uncond_image_clip_feat, image_clip_feat])
    for t in sched.set_timesteps(num_inference_steps):
        x = build_unet_input(latents, first_image_latents, first_images_mask, use_concat)
        pred = unet3d_forward(unet_sd, unet_cfg, x, t, text_embeddings,
                              fps_tensor=dup(fps_tensor), flow_control=dup(flow_control),
                              reference_images_clip_feat=clip,
                              camera_movement_type_tensor=dup(camera_movement_type))
        if video_scale > 0:
            b, f = latents.shape[0], latents.shape[2]
            xs = x.permute(0, 2, 1, 3, 4).reshape(-1, x.shape[1], x.shape[3], x.shape[4]).unsqueeze(2).chunk(2, dim=0)[0]     # :742,745
            ts = torch.cat([text_embeddings] * f, dim=0).chunk(2, dim=0)[0]                                                   # :743,746
            single = unet3d_forward(unet_sd, unet_cfg, xs, t, ts)                                                               # :747-751
            single = single.squeeze(2).reshape(b, f, *single.shape[1:2], *single.shape[3:]).permute(0, 2, 1, 3, 4)            # :755
            u, c = pred.chunk(2)
            noise = single + video_scale * (u - single) + guidance_scale * (c - u)                                            # :757-761
            latents = sched.step(noise, t, latents)
        else:
            latents = sched.step(cfg_combine(pred, guidance_scale), t, latents)
        if trace is not None:
            trace.append(latents.clone())
    return latents


def sample_video(unet_sd, unet_cfg, vae_sd, vae_cfg, sched_cfg, latents, text_embeddings, **kw):
    """Full hot path: denoise then decode; returns (b, 3, f, H, W) fp32 in [0, 1]."""
    return decode_latents(vae_sd, vae_cfg, denoise(unet_sd, unet_cfg, sched_cfg, latents, text_embeddings, **kw))
