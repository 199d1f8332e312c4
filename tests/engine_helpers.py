"""Shared helpers for the GPU parity tests, smoke() and bench.py's checker legs.

Build the product models at "mini" size with the deterministic synthetic state dict, run them on cuda:0 and run the
CPU oracle on the same inputs.  (Imports oracle/: test infrastructure only.)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from followyourclick_b200 import AnimationPipeline, AutoencoderKL, DDIMScheduler, ImageProjModel, UNet3DConditionModel  # noqa: E402
from followyourclick_b200.synth import synth_clip_inputs, synth_state_dict  # noqa: E402
from oracle import ref_pipeline, ref_unet, ref_vae  # noqa: E402
from tests.cfgs import (CLIP_DIM, MINI_VAE, SCHED_V, mini_unet_oracle_cfg, mini_unet_ref_kwargs, unet_inputs)  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLD, name))


def load_synth(model, seed=0):
    sd = model.state_dict()
    new = synth_state_dict({k: tuple(v.shape) for k, v in sd.items()}, seed)
    missing, unexpected = model.load_state_dict(new, strict=False)
    assert not unexpected and all(k.endswith(".pe") for k in missing), (missing, unexpected)
    return {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}


def make_unet(variant, dtype=torch.float32, device="cuda"):
    kw = mini_unet_ref_kwargs(variant)
    unet = UNet3DConditionModel(**kw)
    if kw.get("use_ip_cross_attention"):
        unet.image_proj_model = ImageProjModel(cross_attention_dim=768, clip_embeddings_dim=CLIP_DIM,
                                               clip_extra_context_tokens=kw["num_tokens"])
    sd = load_synth(unet)
    if device is not None:
        unet.to(device)
        unet.to(dtype)
    return unet, sd


def make_vae(dtype=torch.float32, device="cuda", cfg=MINI_VAE):
    vae = AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                        up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=cfg["block_out_channels"],
                        layers_per_block=cfg["layers_per_block"], latent_channels=4, norm_num_groups=32)
    sd = load_synth(vae)
    if device is not None:
        vae.to(device)
        vae.to(dtype)
    return vae, sd


def unet_forward_kwargs(variant, inp, device):
    ocfg = mini_unet_oracle_cfg(variant)
    mv = lambda t: None if t is None else t.to(device)
    return dict(encoder_hidden_states=mv(inp["ctx"]), use_ip_cross_attention=ocfg["use_ip_cross_attention"],
                reference_images_clip_feat=mv(inp.get("clip")), use_camera_motion_condition=ocfg["use_camera_motion_condition"],
                camera_movement_type_tensor=mv(inp.get("camera")), use_fps_condition=ocfg["use_fps_condition"],
                fps_tensor=mv(inp.get("fps")), flow_control=mv(inp.get("flow")))


def stats(a, b):
    """error summary of a (test) against b (reference)"""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    d = (a - b)
    return dict(maxabs=float(d.abs().max()), rel_l2=float(d.norm() / (b.norm() + 1e-12)), ref_max=float(b.abs().max()),
                finite=bool(torch.isfinite(a).all()))


def graph_bookkeeping_on_cpu():
    """True when a CPU test has switched _GraphedUNetStep to its capture-free mode (replay() re-executes the forward): the pipeline's
    graph branch - static input buffers and their views, the graph cache, ClipContext refresh - then runs on CPU too."""
    from followyourclick_b200.pipeline_animation import _GraphedUNetStep
    return not _GraphedUNetStep.capture


def _sync(device):
    if str(device).startswith("cuda"):
        torch.cuda.synchronize()


def run_unet_case(variant, dtype, against="golden", device="cuda"):
    unet, sd = make_unet(variant, dtype, device)
    inp = unet_inputs(variant)
    out = unet(inp["sample"].to(device), inp["timestep"], **unet_forward_kwargs(variant, inp, device)).sample
    _sync(device)
    if against == "golden":
        ref = torch.from_numpy(golden(f"unet_{variant}.npz")["out"])
    else:
        ref = ref_unet.unet3d_forward(sd, mini_unet_oracle_cfg(variant), inp["sample"], inp["timestep"], inp["ctx"],
                                      fps_tensor=inp.get("fps"), flow_control=inp.get("flow"),
                                      reference_images_clip_feat=inp.get("clip"), camera_movement_type_tensor=inp.get("camera"))
    return stats(out, ref)


def run_vae_case(dtype, device="cuda"):
    vae, sd = make_vae(dtype, device)
    g = golden("vae.npz")
    out = vae.decode(torch.from_numpy(g["z"]).to(device)).sample
    _sync(device)
    return stats(out, torch.from_numpy(g["out"]))


def run_vae_encode_case(dtype, device="cuda"):
    """AutoencoderKL.encode against the reference's moments / reparameterised sample (golden vae_encode.npz)."""
    vae, sd = make_vae(dtype, device)
    g = golden("vae_encode.npz")
    dist = vae.encode(torch.from_numpy(g["x"]).to(device)).latent_dist
    noise = torch.from_numpy(g["noise"]).to(device)
    sample = dist.mean + dist.std * noise
    _sync(device)
    return stats(dist.parameters, torch.from_numpy(g["moments"])), stats(sample, torch.from_numpy(g["sample"])), dist


class FakeTokenizer:
    model_max_length = 77

    def __call__(self, prompt, **kw):
        n = len(prompt) if isinstance(prompt, list) else 1
        ids = torch.zeros(n, 77, dtype=torch.long)
        return type("Tok", (), dict(input_ids=ids, attention_mask=torch.ones_like(ids)))()


class FakeTextEncoder(torch.nn.Module):
    """Seeded embeddings stand in for CLIP (outside the hot path): cond for the prompt call, uncond for the negative."""

    def __init__(self, emb):
        super().__init__()
        self.emb, self.calls, self.config = emb, 0, type("Cfg", (), {})()

    def forward(self, ids, attention_mask=None):
        i = self.calls
        self.calls += 1
        e = self.emb[1:2] if i % 2 == 0 else self.emb[0:1]
        return (e.to(ids.device),)


def make_pipeline(dtype, variant="base", vae_cfg=MINI_VAE, device="cuda", clip_inputs=None):
    unet, usd = make_unet(variant, dtype, device)
    vae, vsd = make_vae(dtype, device, vae_cfg)
    ci = clip_inputs if clip_inputs is not None else synth_clip_inputs(1, 4, 8, 8)
    sched = DDIMScheduler(**{k: v for k, v in SCHED_V.items()})
    pipe = AnimationPipeline(vae=vae, text_encoder=FakeTextEncoder(ci["text_embeddings"]), tokenizer=FakeTokenizer(),
                             unet=unet, scheduler=sched)
    pipe.set_progress_bar_config(disable=True)
    return pipe, ci, usd, vsd


def pipeline_call(pipe, ci, F, h, w, steps, gs, **extra):
    pipe.text_encoder.calls = 0
    return pipe("p", negative_prompt="n", video_length=F, height=h * 8, width=w * 8, num_inference_steps=steps,
                guidance_scale=gs, latents=ci["latents"].clone(), use_first_frame_mask_condition_concat=True,
                first_image_latents=ci["first_image_latents"], use_fps_condition=True, fps_tensor=torch.tensor([2]),
                flow_control=torch.tensor([4]), first_images_mask=ci["first_images_mask"], **extra).videos


def run_video_scale_case(dtype, device="cuda", graph=True):
    """per-frame guidance branch (video_scale > 0) against the reference fixture pipeline_video_scale.npz."""
    g = golden("pipeline_video_scale.npz")
    pipe, ci, usd, vsd = make_pipeline(dtype, device=device)
    pipe.use_cuda_graph = graph and (str(device).startswith("cuda") or graph_bookkeeping_on_cpu())
    video = pipeline_call(pipe, ci, 4, 8, 8, int(g["steps"]), 8.0, video_scale=float(g["video_scale"]))
    ref = torch.from_numpy(g["video"])
    s = stats(video, ref)
    mse = float(((video.float() - ref) ** 2).mean())
    return dict(video_maxabs=s["maxabs"], psnr=float(10 * np.log10(1.0 / max(mse, 1e-20))), finite=s["finite"], shape=tuple(video.shape))


def run_pipeline_case(dtype, steps=3, against="oracle", device="cuda"):
    """cfg1-style plumbing at mini size: F=4, 8x8 latents, CFG 8.0, mask/first-frame concat, fps/flow condition."""
    F, h, w, gs = 4, 8, 8, 8.0
    pipe, ci, usd, vsd = make_pipeline(dtype, device=device)
    if not str(device).startswith("cuda"):
        pipe.use_cuda_graph = graph_bookkeeping_on_cpu()
    video = pipeline_call(pipe, ci, F, h, w, steps, gs)
    if against == "golden":
        assert steps == 3
        ref = torch.from_numpy(golden("pipeline.npz")["video"])
    else:
        lat = ref_pipeline.denoise(usd, mini_unet_oracle_cfg("base"), SCHED_V, ci["latents"], ci["text_embeddings"], steps, gs,
                                   first_image_latents=ci["first_image_latents"], first_images_mask=ci["first_images_mask"],
                                   fps_tensor=torch.tensor([2]), flow_control=torch.tensor([4]))
        ref = ref_vae.decode_latents(vsd, MINI_VAE, lat)
    s = stats(video, ref)
    mse = float(((video.float() - ref) ** 2).mean())
    return dict(video_maxabs=s["maxabs"], video_rel_l2=s["rel_l2"], psnr=float(10 * np.log10(1.0 / max(mse, 1e-20))),
                shape=tuple(video.shape), finite=s["finite"])


def run_unet2d_case(dtype, device="cuda"):
    """UNet2DConditionModel (T2I first-frame generator) against the reference's 2-D UNet output (unet2d.npz);
    also checks the state-dict key set is the 2-D checkpoint's."""
    import json
    from followyourclick_b200 import UNet2DConditionModel
    from tests.cfgs import MINI_UNET2D
    m = UNet2DConditionModel(**MINI_UNET2D)
    keys = json.load(open(os.path.join(GOLD, "unet2d_keys.json")))
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == keys
    load_synth(m)
    m.to(device)
    m.to(dtype)
    g = golden("unet2d.npz")
    out = m(torch.from_numpy(g["x"]).to(device), torch.tensor(501), encoder_hidden_states=torch.from_numpy(g["ctx"]).to(device)).sample
    _sync(device)
    assert out.shape == (2, 4, 16, 16)
    return stats(out, torch.from_numpy(g["out"]))


def run_unet_ragged_case(dtype, device="cuda", b=1, f=3, h=24, w=40):
    """Non-square, non-power-of-two latent grid, odd frame count, batch 1 (no CFG pair): the shapes where tile pickers fall back
    (tcgen05 self-attention needs L % 128 == 0, conv patches need power-of-two factors).  Checked against the oracle run in place."""
    from tests.cfgs import unet_inputs
    unet, sd = make_unet("base", dtype, device)
    inp = unet_inputs("base", b=b, f=f, h=h, w=w, seed=23)
    inp["fps"], inp["flow"] = torch.tensor([2] * b), torch.tensor([4] * b)
    out = unet(inp["sample"].to(device), inp["timestep"], **unet_forward_kwargs("base", inp, device)).sample
    _sync(device)
    if (b, f, h, w) == (1, 3, 24, 40):          # this shape has a fixture from the unmodified reference (tests/golden/unet_base_ragged.npz)
        ref = torch.from_numpy(golden("unet_base_ragged.npz")["out"])
    else:
        ref = ref_unet.unet3d_forward(sd, mini_unet_oracle_cfg("base"), inp["sample"], inp["timestep"], inp["ctx"],
                                      fps_tensor=inp["fps"], flow_control=inp["flow"])
    assert out.shape == ref.shape == (b, 4, f, h, w)
    return stats(out, ref)


class FakeIPAdapter:
    """MyIPAdapter stand-in at pipeline level: seeded (cond, uncond) CLIP features (the vision tower is outside the hot path)."""

    def __init__(self, cond, uncond):
        self.cond, self.uncond = cond, uncond

    def get_image_clip_feat(self, input_image=None):
        return self.cond, self.uncond


def run_pipeline_variant_case(variant, dtype, device="cuda", graph=True):
    """BASELINE configs[2] ('ip': shipped YAML + IP-Adapter image condition) and configs[4] ('cam': camera-LoRA model, epsilon prediction,
    4-channel input) plumbing at mini size against the UNMODIFIED reference pipeline's frames (tests/golden/pipeline_{variant}.npz)."""
    from tests.cfgs import pipeline_variant_inputs
    ci, kw, _, sched_cfg, steps, gs = pipeline_variant_inputs(variant)
    unet, _ = make_unet(variant, dtype, device)
    vae, _ = make_vae(dtype, device)
    pipe = AnimationPipeline(vae=vae, text_encoder=FakeTextEncoder(ci["text_embeddings"]), tokenizer=FakeTokenizer(), unet=unet,
                             scheduler=DDIMScheduler(**sched_cfg),
                             ip_adapter=FakeIPAdapter(ci["image_clip_feat"].to(device), ci["uncond_image_clip_feat"].to(device)))
    pipe.set_progress_bar_config(disable=True)
    pipe.use_cuda_graph = graph and (str(device).startswith("cuda") or graph_bookkeeping_on_cpu())
    video = pipe("p", negative_prompt="n", video_length=4, height=64, width=64, num_inference_steps=steps, guidance_scale=gs,
                 latents=ci["latents"].clone(), use_ip_cross_attention=True, condition_images=torch.zeros(1, 3, 8, 8), **kw).videos
    ref = torch.from_numpy(golden(f"pipeline_{variant}.npz")["video"])
    s = stats(video, ref)
    mse = float(((video.float() - ref) ** 2).mean())
    return dict(video_maxabs=s["maxabs"], psnr=float(10 * np.log10(1.0 / max(mse, 1e-20))), finite=s["finite"], shape=tuple(video.shape))


def run_ip_attn_processor_case(name, dtype, device="cuda"):
    """followyourclick_b200.ip_adapter.IPAttnProcessor (the north star's named call surface) against the output of the
    UNMODIFIED reference processor (tests/golden/ip_attn_processor.npz, ip_adapter/attention_processor.py:80-183) on the same duck-typed
    ``attn`` module, weights and inputs."""
    from followyourclick_b200.ip_adapter import IPAttnProcessor
    from tests.cfgs import IP_ATTN_SCALE, DuckAttention, ip_attn_case
    c = ip_attn_case(name)
    attn = DuckAttention(c).to(device)
    proc = IPAttnProcessor(hidden_size=c["C"], cross_attention_dim=c["xd"], scale=IP_ATTN_SCALE, num_tokens=c["T"]).to(device)
    with torch.no_grad():
        proc.to_k_ip.weight.copy_(c["w"]["to_k_ip"]); proc.to_v_ip.weight.copy_(c["w"]["to_v_ip"])
    assert sorted(proc.state_dict()) == ["to_k_ip.weight", "to_v_ip.weight"]           # the adapter checkpoint's key names
    y = proc(attn, c["x"].to(device=device, dtype=dtype), encoder_hidden_states=c["ctx"].to(device=device, dtype=dtype))
    _sync(device)
    ref = torch.from_numpy(golden("ip_attn_processor.npz")[name])
    assert y.shape == ref.shape and y.dtype == dtype
    last = y.float().cpu()
    if c["shape4d"]:                       # channel-last for the per-row statistics
        last, ref = last.permute(0, 2, 3, 1), ref.permute(0, 2, 3, 1)
    return stats(last, ref)
