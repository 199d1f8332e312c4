"""Generate the golden fixtures from the UNMODIFIED reference (build container only).

Run:  python tests/golden/make_golden.py        (needs /root/reference; ~1 min on 8 cores)

What it does, per fixture:
  1. imports the reference (animatediff/, diffusers 0.11.1 vendored copy) with the three import shims of
     SURVEY App. C - nothing is copied into this repo;
  2. builds the reference model at a reduced ("mini") size, overwrites every tensor with the deterministic
     synthetic weights of followyourclick_b200/synth.py (key-name seeded, so the product and the oracle can
     regenerate the identical state dict without shipping it);
  3. runs the reference on seeded inputs, checks that oracle/ reproduces the reference output (the oracle is
     *pinned* here: max-abs error is recorded in the fixture), and stores inputs+outputs as small .npz files.

The fixtures travel to the GPU box; /root/reference does not.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference():
    sys.path.insert(0, REF)
    import huggingface_hub as hh
    for n in ("HfFolder", "cached_download"):
        if not hasattr(hh, n):
            setattr(hh, n, object)
    pkg = types.ModuleType("diffusers")
    pkg.__path__ = [REF + "/diffusers"]
    pkg.__version__ = "0.11.1"
    sys.modules["diffusers"] = pkg
    sys.modules["diffusers.pipelines"] = types.ModuleType("diffusers.pipelines")
    pkg.pipelines = sys.modules["diffusers.pipelines"]
    sys.modules["imageio"] = types.ModuleType("imageio")
    pkg.StableDiffusionPipeline = object              # only imported by ip_adapter/my_ip_adapter.py:5, never used here
    ipa = types.ModuleType("ip_adapter")              # skip ip_adapter/__init__.py (pulls modern-diffusers pipelines)
    ipa.__path__ = [REF + "/ip_adapter"]
    sys.modules["ip_adapter"] = ipa
    from animatediff.models.unet import UNet3DConditionModel
    from animatediff.pipelines.pipeline_animation import AnimationPipeline
    from diffusers.models.vae import AutoencoderKL
    from diffusers.schedulers.scheduling_ddim import DDIMScheduler
    from ip_adapter.my_ip_adapter import ImageProjModel
    return UNet3DConditionModel, AnimationPipeline, AutoencoderKL, DDIMScheduler, ImageProjModel


from tests.cfgs import (MINI_UNET_VARIANTS, MINI_VAE, SCHED_EPS, SCHED_V, mini_unet_ref_kwargs, mini_unet_oracle_cfg,
                        unet_inputs, CLIP_DIM)  # noqa: E402
from followyourclick_b200.synth import synth_state_dict, synth_clip_inputs  # noqa: E402
from oracle import ref_ddim, ref_pipeline, ref_unet, ref_vae  # noqa: E402


def load_synth(model, seed=0):
    sd = model.state_dict()
    new = synth_state_dict({k: tuple(v.shape) for k, v in sd.items()}, seed)
    missing, unexpected = model.load_state_dict(new, strict=False)
    assert not unexpected and all(k.endswith(".pe") for k in missing), (missing, unexpected)
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def maxabs(a, b):
    return float((a - b).abs().max())


def vae_encode_fixture(VAE):
    """encoder moments of the unmodified reference on a seeded image, oracle checked against them."""
    from tests.cfgs import MINI_VAE
    vae = VAE(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
              up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=MINI_VAE["block_out_channels"],
              layers_per_block=MINI_VAE["layers_per_block"], latent_channels=4, norm_num_groups=32).eval()
    vsd = load_synth(vae)
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(11)) * 2 - 1
    noise = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(12))
    with torch.no_grad():
        dist = vae.encode(x).latent_dist
        ref_m = dist.parameters
        ref_s = dist.mean + dist.std * noise
        orc_m = ref_vae.vae_encode_moments(vsd, MINI_VAE, x)
        orc_s = ref_vae.gaussian_sample(orc_m, noise)
    err = max(maxabs(ref_m, orc_m), maxabs(ref_s, orc_s))
    print(f"vae encode moments {tuple(ref_m.shape)} |ref|max={float(ref_m.abs().max()):.3f} oracle-vs-ref maxabs={err:.3e}")
    assert err < 2e-4 * max(1.0, float(ref_m.abs().max()))
    np.savez_compressed(os.path.join(HERE, "vae_encode.npz"), x=x.numpy(), noise=noise.numpy(), moments=ref_m.numpy(), sample=ref_s.numpy())
    return err


def resampler_fixture():
    """the unmodified reference Resampler (ip_adapter/resampler.py) on seeded CLIP-like features."""
    from ip_adapter.resampler import Resampler
    from oracle import ref_resampler
    from tests.cfgs import MINI_RESAMPLER, RESAMPLER_TOKENS
    m = Resampler(**MINI_RESAMPLER).eval()
    sd = load_synth(m)
    assert {k: tuple(v.shape) for k, v in sd.items()} == ref_resampler.resampler_param_shapes(MINI_RESAMPLER)
    x = torch.randn(2, RESAMPLER_TOKENS, MINI_RESAMPLER["embedding_dim"], generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        ref = m(x)
        orc = ref_resampler.resampler_forward(sd, MINI_RESAMPLER, x)
    err = maxabs(ref, orc)
    print(f"resampler out {tuple(ref.shape)} |ref|max={float(ref.abs().max()):.3f} oracle-vs-ref maxabs={err:.3e}")
    assert err < 2e-5
    np.savez_compressed(os.path.join(HERE, "resampler.npz"), x=x.numpy(), out=ref.numpy())
    return err


class FakeTok:
    model_max_length = 77

    def __call__(self, prompt, **kw):
        n = len(prompt) if isinstance(prompt, list) else 1
        ids = torch.zeros(n, 77, dtype=torch.long)
        return types.SimpleNamespace(input_ids=ids, attention_mask=torch.ones_like(ids))

    def batch_decode(self, x):
        return [""]


class FakeText(torch.nn.Module):
    """Returns the seeded 'cond' embedding for the prompt call and 'uncond' for the negative-prompt call."""
    def __init__(self, emb):
        super().__init__()
        self.emb, self.calls = emb, 0
        self.config = types.SimpleNamespace()
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, ids, attention_mask=None):
        i = self.calls
        self.calls += 1
        return (self.emb[1:2] if i % 2 == 0 else self.emb[0:1],)     # prompt first, then negative prompt


def video_scale_fixture(UNet, Pipe, VAE, DDIM):
    """the reference pipeline with video_scale > 0 (per-frame guidance branch, pipeline_animation.py:738-761)."""
    vae = VAE(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
              up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=MINI_VAE["block_out_channels"],
              layers_per_block=MINI_VAE["layers_per_block"], latent_channels=4, norm_num_groups=32).eval()
    vsd = load_synth(vae)
    unet = UNet(**mini_unet_ref_kwargs("base")).eval()
    usd = load_synth(unet)
    F_, h, w, steps, gs, vs = 4, 8, 8, 2, 8.0, 0.7
    ci = synth_clip_inputs(1, F_, h, w)
    sched = DDIM(**{k: v for k, v in SCHED_V.items() if k != "set_alpha_to_one"})
    pipe = Pipe(vae=vae, text_encoder=FakeText(ci["text_embeddings"]), tokenizer=FakeTok(), unet=unet, scheduler=sched)
    with torch.no_grad():
        ref_video = pipe("p", negative_prompt="n", video_length=F_, height=h * 8, width=w * 8,
                         num_inference_steps=steps, guidance_scale=gs, latents=ci["latents"].clone(),
                         use_first_frame_mask_condition_concat=True, first_image_latents=ci["first_image_latents"],
                         use_fps_condition=True, fps_tensor=torch.tensor([2]), flow_control=torch.tensor([4]),
                         first_images_mask=ci["first_images_mask"], video_scale=vs).videos
        lat = ref_pipeline.denoise(usd, mini_unet_oracle_cfg("base"), SCHED_V, ci["latents"], ci["text_embeddings"],
                                   steps, gs, first_image_latents=ci["first_image_latents"],
                                   first_images_mask=ci["first_images_mask"], fps_tensor=torch.tensor([2]),
                                   flow_control=torch.tensor([4]), video_scale=vs)
        orc_video = ref_vae.decode_latents(vsd, MINI_VAE, lat)
    err = maxabs(ref_video, orc_video)
    print(f"pipeline video_scale={vs} video {tuple(ref_video.shape)} oracle-vs-ref maxabs={err:.3e}")
    assert err < 2e-3
    np.savez_compressed(os.path.join(HERE, "pipeline_video_scale.npz"), video=ref_video.numpy().astype(np.float32),
                        final_latents=lat.numpy(), video_scale=np.float32(vs), steps=np.int64(steps))
    return err


def unet2d_fixture():
    """the stock 2-D UNet (T2I first-frame generator) of the vendored diffusers, mini size, synthetic weights;
    the oracle is the 3-D restatement without motion modules on one frame."""
    from diffusers.models.unet_2d_condition import UNet2DConditionModel as UNet2D
    from tests.cfgs import MINI_UNET2D
    m = UNet2D(**MINI_UNET2D).eval()
    sd = load_synth(m)
    g = torch.Generator().manual_seed(31)
    x, ctx, t = torch.randn(2, 4, 16, 16, generator=g), torch.randn(2, 77, 768, generator=g), torch.tensor(501)
    from tests.cfgs import mini_unet2d_oracle_cfg
    with torch.no_grad():
        ref = m(x, t, encoder_hidden_states=ctx).sample
        orc = ref_unet.unet3d_forward(sd, mini_unet2d_oracle_cfg(), x.unsqueeze(2), t, ctx).squeeze(2)
    err = maxabs(ref, orc)
    print(f"unet2d out {tuple(ref.shape)} |ref|max={float(ref.abs().max()):.3f} oracle-vs-ref maxabs={err:.3e}")
    assert err < 2e-4 * max(1.0, float(ref.abs().max()))
    np.savez_compressed(os.path.join(HERE, "unet2d.npz"), x=x.numpy(), ctx=ctx.numpy(), out=ref.numpy())
    with open(os.path.join(HERE, "unet2d_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in sd.items()}, f)
    return err


class FakeIP:
    """Stands in for MyIPAdapter at pipeline level (pipeline_animation.py:676-680): returns seeded (cond, uncond) CLIP features."""
    def __init__(self, cond, uncond):
        self.cond, self.uncond = cond, uncond

    def get_image_clip_feat(self, input_image=None):
        return self.cond, self.uncond


def pipeline_variant_fixture(UNet, Pipe, VAE, DDIM, ImageProjModel, variant):
    """BASELINE configs[2] / [4] plumbing at mini size through the UNMODIFIED reference pipeline: 'ip' = shipped YAML + IP-Adapter
    image condition ([uncond, cond] CLIP features, 9-channel input, v-prediction); 'cam' = the camera-LoRA model of
    inference_w_camera_lora.py (4-channel input, epsilon prediction, camera-motion embedding, temporal LoRA, IP tokens)."""
    from tests.cfgs import pipeline_variant_inputs
    vae = VAE(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
              up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=MINI_VAE["block_out_channels"],
              layers_per_block=MINI_VAE["layers_per_block"], latent_channels=4, norm_num_groups=32).eval()
    vsd = load_synth(vae)
    unet = UNet(**mini_unet_ref_kwargs(variant)).eval()
    ocfg = mini_unet_oracle_cfg(variant)
    unet.image_proj_model = ImageProjModel(cross_attention_dim=768, clip_embeddings_dim=CLIP_DIM, clip_extra_context_tokens=ocfg["num_tokens"])
    usd = load_synth(unet)
    ci, kw, okw, sched_cfg, steps, gs = pipeline_variant_inputs(variant)
    sched = DDIM(**{k: v for k, v in sched_cfg.items() if k != "set_alpha_to_one"})
    pipe = Pipe(vae=vae, text_encoder=FakeText(ci["text_embeddings"]), tokenizer=FakeTok(), unet=unet, scheduler=sched,
                ip_adapter=FakeIP(ci["image_clip_feat"], ci["uncond_image_clip_feat"]))
    with torch.no_grad():
        ref_video = pipe("p", negative_prompt="n", video_length=4, height=64, width=64, num_inference_steps=steps, guidance_scale=gs,
                         latents=ci["latents"].clone(), use_ip_cross_attention=True, condition_images=torch.zeros(1, 3, 8, 8), **kw).videos
        lat = ref_pipeline.denoise(usd, ocfg, sched_cfg, ci["latents"], ci["text_embeddings"], steps, gs,
                                   image_clip_feat=ci["image_clip_feat"], uncond_image_clip_feat=ci["uncond_image_clip_feat"], **okw)
        orc_video = ref_vae.decode_latents(vsd, MINI_VAE, lat)
    err = maxabs(ref_video, orc_video)
    print(f"pipeline[{variant}] video {tuple(ref_video.shape)} oracle-vs-ref maxabs={err:.3e}")
    assert err < 2e-3
    np.savez_compressed(os.path.join(HERE, f"pipeline_{variant}.npz"), video=ref_video.numpy().astype(np.float32), final_latents=lat.numpy())
    return err


def ip_attn_processor_fixture():
    """the reference's ``IPAttnProcessor.__call__`` (ip_adapter/attention_processor.py:80-183), loaded from its
    file (ip_adapter/__init__.py pulls modern-diffusers pipelines) and driven with a duck-typed ``attn`` (tests/cfgs.DuckAttention:
    the helpers of the modern pip-diffusers ``Attention`` class the processor targets, absent from /root/reference).  Cases
    (tests/cfgs.IP_ATTN_CASES): 3-D tokens with T = 4, 4-D (b, c, h, w) input with T = 16, residual_connection=True.  Only the
    reference outputs are stored; the oracle's two-softmax formula is checked against them here."""
    import importlib.util
    from tests.cfgs import IP_ATTN_CASES, IP_ATTN_SCALE, DuckAttention, ip_attn_case
    spec = importlib.util.spec_from_file_location("_ref_attention_processor", REF + "/ip_adapter/attention_processor.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out, worst = {}, 0.0
    for name in IP_ATTN_CASES:
        c = ip_attn_case(name)
        C, heads, T, w, x, ctx = c["C"], c["heads"], c["T"], c["w"], c["x"], c["ctx"]
        attn = DuckAttention(c)
        proc = mod.IPAttnProcessor(hidden_size=C, cross_attention_dim=c["xd"], scale=IP_ATTN_SCALE, num_tokens=T)
        with torch.no_grad():
            proc.to_k_ip.weight.copy_(w["to_k_ip"]); proc.to_v_ip.weight.copy_(w["to_v_ip"])
            y = proc(attn, x, encoder_hidden_states=ctx)
            tok = x.reshape(x.shape[0], C, -1).transpose(1, 2) if c["shape4d"] else x
            q = tok @ w["to_q"].t()
            o = ref_unet.mha(q, ctx[:, :-T] @ w["to_k"].t(), ctx[:, :-T] @ w["to_v"].t(), heads) \
                + IP_ATTN_SCALE * ref_unet.mha(q, ctx[:, -T:] @ w["to_k_ip"].t(), ctx[:, -T:] @ w["to_v_ip"].t(), heads)
            yo = o @ w["to_out_w"].t() + w["to_out_b"]
            if c["residual"]:
                yo = yo + tok
            if c["shape4d"]:
                yo = yo.transpose(1, 2).reshape(x.shape)
        worst = max(worst, maxabs(y, yo))
        out[name] = y.numpy()
    print(f"IPAttnProcessor fixture: oracle-vs-ref maxabs={worst:.3e}")
    assert worst < 2e-5
    np.savez_compressed(os.path.join(HERE, "ip_attn_processor.npz"), **out)
    return worst


def unet_ragged_fixture(UNet):
    """The unmodified reference UNet on a non-square, non-power-of-two latent grid with an odd frame count and batch 1 (320 x 192
    image, 3 frames): pins the oracle (and through it the engine's tile-picker fallbacks) away from the 16 x 16 x 4 fixture shape."""
    unet = UNet(**mini_unet_ref_kwargs("base")).eval()
    sd = load_synth(unet)
    inp = unet_inputs("base", b=1, f=3, h=24, w=40, seed=23)
    fps, flow = torch.tensor([2]), torch.tensor([4])
    with torch.no_grad():
        ref = unet(inp["sample"], inp["timestep"], encoder_hidden_states=inp["ctx"], use_fps_condition=True, fps_tensor=fps, flow_control=flow).sample
        orc = ref_unet.unet3d_forward(sd, mini_unet_oracle_cfg("base"), inp["sample"], inp["timestep"], inp["ctx"], fps_tensor=fps, flow_control=flow)
    err = maxabs(ref, orc)
    print(f"unet[base, ragged 1x3x24x40] |ref|max={float(ref.abs().max()):.3f} oracle-vs-ref maxabs={err:.3e}")
    assert err < 2e-4 * max(1.0, float(ref.abs().max()))
    np.savez_compressed(os.path.join(HERE, "unet_base_ragged.npz"), out=ref.numpy())
    return err


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    UNet, Pipe, VAE, DDIM, ImageProjModel = import_reference()
    if "--only-unet-ragged" in sys.argv:
        pj = os.path.join(HERE, "pins.json")
        d = json.load(open(pj))
        d["oracle_vs_reference_maxabs"]["unet_base_ragged"] = unet_ragged_fixture(UNet)
        json.dump(d, open(pj, "w"), indent=1)
        return
    if "--only-pipeline-variants" in sys.argv:
        pj = os.path.join(HERE, "pins.json")
        d = json.load(open(pj))
        for v in ("ip", "cam"):
            d["oracle_vs_reference_maxabs"][f"pipeline_{v}"] = pipeline_variant_fixture(UNet, Pipe, VAE, DDIM, ImageProjModel, v)
        json.dump(d, open(pj, "w"), indent=1)
        return
    if "--only-ip-attn-processor" in sys.argv:
        pj = os.path.join(HERE, "pins.json")
        d = json.load(open(pj))
        d["oracle_vs_reference_maxabs"]["ip_attn_processor"] = ip_attn_processor_fixture()
        json.dump(d, open(pj, "w"), indent=1)
        return
    if "--only-unet2d" in sys.argv:
        pj = os.path.join(HERE, "pins.json")
        d = json.load(open(pj))
        d["oracle_vs_reference_maxabs"]["unet2d"] = unet2d_fixture()
        json.dump(d, open(pj, "w"), indent=1)
        return
    if "--only-video-scale" in sys.argv:
        pj = os.path.join(HERE, "pins.json")
        d = json.load(open(pj))
        d["oracle_vs_reference_maxabs"]["pipeline_video_scale"] = video_scale_fixture(UNet, Pipe, VAE, DDIM)
        json.dump(d, open(pj, "w"), indent=1)
        return
    if "--only-resampler" in sys.argv:           # add the resampler fixture without regenerating the others
        pj = os.path.join(HERE, "pins.json")
        d = json.load(open(pj))
        d["oracle_vs_reference_maxabs"]["resampler"] = resampler_fixture()
        json.dump(d, open(pj, "w"), indent=1)
        return
    if "--only-vae-encode" in sys.argv:          # add the encoder fixture without regenerating the others
        pj = os.path.join(HERE, "pins.json")
        d = json.load(open(pj))
        d["oracle_vs_reference_maxabs"]["vae_encode"] = vae_encode_fixture(VAE)
        json.dump(d, open(pj, "w"), indent=1)
        return
    pins = {}

    # ------------------------------------------------------------------ DDIM known answers
    out = {}
    for name, cfg in (("v", SCHED_V), ("eps", SCHED_EPS)):
        ref = DDIM(**{k: v for k, v in cfg.items() if k != "set_alpha_to_one"})
        orc = ref_ddim.DDIMOracle(cfg)
        assert torch.equal(ref.alphas_cumprod, orc.alphas_cumprod)
        out[f"{name}_alphas_cumprod"] = ref.alphas_cumprod.numpy()
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(1, 4, 4, 8, 8, generator=g)
        v = torch.randn(1, 4, 4, 8, 8, generator=g)
        out[f"{name}_x"], out[f"{name}_v"] = x.numpy(), v.numpy()
        for n in (4, 25, 50):
            ref.set_timesteps(n)
            ts = orc.set_timesteps(n)
            assert torch.equal(ref.timesteps, ts)
            out[f"{name}_timesteps_{n}"] = ts.numpy()
            for t in (int(ts[0]), int(ts[len(ts) // 2]), int(ts[-1])):
                r = ref.step(v, t, x).prev_sample
                o = orc.step(v, t, x)
                assert torch.equal(r, o), (name, n, t, maxabs(r, o))
                out[f"{name}_step_{n}_{t}"] = r.numpy()
        # eta > 0 with explicit variance noise
        ref.set_timesteps(25); orc.set_timesteps(25)
        noise = torch.randn(1, 4, 4, 8, 8, generator=g)
        r = ref.step(v, 481, x, eta=0.5, variance_noise=noise).prev_sample
        o = orc.step(v, 481, x, eta=0.5, variance_noise=noise)
        assert torch.equal(r, o)
        out[f"{name}_noise"] = noise.numpy()
        out[f"{name}_step_eta0.5_25_481"] = r.numpy()
    np.savez_compressed(os.path.join(HERE, "ddim.npz"), **out)
    pins["ddim"] = 0.0
    print("ddim ok (oracle bit-exact vs reference)")

    # ------------------------------------------------------------------ UNet3D forward, mini size
    keyfile = {}
    for variant in MINI_UNET_VARIANTS:
        kw = mini_unet_ref_kwargs(variant)
        unet = UNet(**kw).eval()
        ocfg = mini_unet_oracle_cfg(variant)
        if ocfg["use_ip_cross_attention"]:
            unet.image_proj_model = ImageProjModel(cross_attention_dim=768, clip_embeddings_dim=CLIP_DIM,
                                                   clip_extra_context_tokens=ocfg["num_tokens"])
        sd = load_synth(unet)
        keyfile[variant] = {k: list(v.shape) for k, v in sd.items()}
        inp = unet_inputs(variant)
        with torch.no_grad():
            ref_out = unet(inp["sample"], inp["timestep"], encoder_hidden_states=inp["ctx"],
                           use_ip_cross_attention=ocfg["use_ip_cross_attention"],
                           reference_images_clip_feat=inp.get("clip"),
                           use_camera_motion_condition=ocfg["use_camera_motion_condition"],
                           camera_movement_type_tensor=inp.get("camera"),
                           use_fps_condition=ocfg["use_fps_condition"],
                           fps_tensor=inp.get("fps"), flow_control=inp.get("flow")).sample
            taps = {}
            orc_out = ref_unet.unet3d_forward(sd, ocfg, inp["sample"], inp["timestep"], inp["ctx"],
                                              fps_tensor=inp.get("fps"), flow_control=inp.get("flow"),
                                              reference_images_clip_feat=inp.get("clip"),
                                              camera_movement_type_tensor=inp.get("camera"), taps=taps)
        err = maxabs(ref_out, orc_out)
        scale = float(ref_out.abs().max())
        print(f"unet[{variant}] out {tuple(ref_out.shape)} |ref|max={scale:.3f} oracle-vs-ref maxabs={err:.3e}")
        assert err < 2e-4 * max(scale, 1.0), err
        pins[f"unet_{variant}"] = err
        np.savez_compressed(os.path.join(HERE, f"unet_{variant}.npz"), out=ref_out.numpy(),
                            **{"tap_" + k: v.numpy().astype(np.float16) for k, v in taps.items() if k in ("conv_in", "mid")})
    with open(os.path.join(HERE, "unet_keys.json"), "w") as f:
        json.dump(keyfile, f)

    # ------------------------------------------------------------------ VAE decode, mini size
    vae = VAE(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
              up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=MINI_VAE["block_out_channels"],
              layers_per_block=MINI_VAE["layers_per_block"], latent_channels=4, norm_num_groups=32).eval()
    vsd = load_synth(vae)
    with open(os.path.join(HERE, "vae_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in vsd.items()}, f)
    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        ref_out = vae.decode(z).sample
        orc_out = ref_vae.vae_decode(vsd, MINI_VAE, z)
    err = maxabs(ref_out, orc_out)
    print(f"vae out {tuple(ref_out.shape)} |ref|max={float(ref_out.abs().max()):.3f} oracle-vs-ref maxabs={err:.3e}")
    assert err < 2e-4 * max(1.0, float(ref_out.abs().max()))
    pins["vae"] = err
    np.savez_compressed(os.path.join(HERE, "vae.npz"), z=z.numpy(), out=ref_out.numpy())

    # ------------------------------------------------------------------ full pipeline, mini size (cfg1-style plumbing)
    class FakeTok:
        model_max_length = 77

        def __call__(self, prompt, **kw):
            n = len(prompt) if isinstance(prompt, list) else 1
            ids = torch.zeros(n, 77, dtype=torch.long)
            return types.SimpleNamespace(input_ids=ids, attention_mask=torch.ones_like(ids))

        def batch_decode(self, x):
            return [""]

    class FakeText(torch.nn.Module):
        """Returns the seeded 'cond' embedding for the prompt call and 'uncond' for the negative-prompt call."""
        def __init__(self, emb):
            super().__init__()
            self.emb, self.calls = emb, 0
            self.config = types.SimpleNamespace()
            self.dummy = torch.nn.Parameter(torch.zeros(1))

        def forward(self, ids, attention_mask=None):
            i = self.calls
            self.calls += 1
            return (self.emb[1:2] if i % 2 == 0 else self.emb[0:1],)     # prompt first, then negative prompt

    variant = "base"
    unet = UNet(**mini_unet_ref_kwargs(variant)).eval()
    usd = load_synth(unet)
    F_, h, w, steps, gs = 4, 8, 8, 3, 8.0
    ci = synth_clip_inputs(1, F_, h, w)
    sched = DDIM(**{k: v for k, v in SCHED_V.items() if k != "set_alpha_to_one"})
    pipe = Pipe(vae=vae, text_encoder=FakeText(ci["text_embeddings"]), tokenizer=FakeTok(), unet=unet, scheduler=sched)
    with torch.no_grad():
        ref_video = pipe("p", negative_prompt="n", video_length=F_, height=h * 8, width=w * 8,
                         num_inference_steps=steps, guidance_scale=gs, latents=ci["latents"].clone(),
                         use_first_frame_mask_condition_concat=True, first_image_latents=ci["first_image_latents"],
                         use_fps_condition=True, fps_tensor=torch.tensor([2]), flow_control=torch.tensor([4]),
                         first_images_mask=ci["first_images_mask"]).videos
        trace = []
        lat = ref_pipeline.denoise(usd, mini_unet_oracle_cfg(variant), SCHED_V, ci["latents"], ci["text_embeddings"],
                                   steps, gs, first_image_latents=ci["first_image_latents"],
                                   first_images_mask=ci["first_images_mask"], fps_tensor=torch.tensor([2]),
                                   flow_control=torch.tensor([4]), trace=trace)
        orc_video = ref_vae.decode_latents(vsd, MINI_VAE, lat)
    err = maxabs(ref_video, orc_video)
    print(f"pipeline video {tuple(ref_video.shape)} oracle-vs-ref maxabs={err:.3e}")
    assert err < 2e-3
    pins["pipeline"] = err
    pins["vae_encode"] = vae_encode_fixture(VAE)
    pins["resampler"] = resampler_fixture()
    pins["unet2d"] = unet2d_fixture()
    pins["unet_base_ragged"] = unet_ragged_fixture(UNet)
    pins["ip_attn_processor"] = ip_attn_processor_fixture()
    for v in ("ip", "cam"):
        pins[f"pipeline_{v}"] = pipeline_variant_fixture(UNet, Pipe, VAE, DDIM, ImageProjModel, v)
    pins["pipeline_video_scale"] = video_scale_fixture(UNet, Pipe, VAE, DDIM)
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), video=ref_video.numpy().astype(np.float32),
                        final_latents=lat.numpy())
    with open(os.path.join(HERE, "pins.json"), "w") as f:
        json.dump({"oracle_vs_reference_maxabs": pins, "torch": torch.__version__,
                   "reference": "mayuelala/FollowYourClick @ /root/reference (unmodified)"}, f, indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
