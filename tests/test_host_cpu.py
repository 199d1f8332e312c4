"""CPU: host-side logic of the drop-in classes and the C-ABI surface (no kernel launches)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from tests.cfgs import MINI_UNET_VARIANTS, MINI_VAE, SCHED_EPS, SCHED_V, mini_unet_ref_kwargs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_loads_and_exports_every_header_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "fyc.h")).read()
    declared = set(re.findall(r"\b(fyc_[a-z0-9_]+)\s*\(", header))
    from followyourclick_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib().fyc_version() == 100
    assert _lib.lib().fyc_last_error() is not None


@pytest.mark.parametrize("variant", MINI_UNET_VARIANTS)
def test_unet_state_dict_key_contract(variant):
    """Same keys and shapes as the reference model built with the same kwargs (fixture written by make_golden.py)."""
    from followyourclick_b200 import UNet3DConditionModel
    ref = {k: tuple(s) for k, s in json.load(open(os.path.join(GOLD, "unet_keys.json")))[variant].items()
           if not k.startswith("image_proj_model.")}
    m = UNet3DConditionModel(**mini_unet_ref_kwargs(variant))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == ref
    assert m.config.cross_attention_dim == 768 and m.config["_diffusers_version"] == "0.11.1" and m.in_channels == 4
    # checkpoints wrapped with extra keys load with strict=False and report them (scripts/inference.py:178-181)
    sd = m.state_dict()
    sd["not.a.key"] = torch.zeros(1)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert list(unexpected) == ["not.a.key"] and not missing
    m.enable_xformers_memory_efficient_attention()
    assert m._xformers_semantics


def test_full_size_key_count_and_pe_buffer():
    """SD-1.5-size model with the shipped inference YAML + IP: 1286 tensors, pe buffers == formula."""
    from followyourclick_b200 import ImageProjModel
    from followyourclick_b200.unet import sinusoidal_pe, unet_param_spec
    from oracle.ref_unet import default_unet_config
    cfg = default_unet_config(use_ip_cross_attention=True, use_first_frame_condition_concat=False, use_camera_motion_condition=False)
    spec = unet_param_spec(cfg)
    assert len(spec) == 1286 and len(ImageProjModel(768, 1024, 4).state_dict()) == 4
    assert spec["conv_in.weight"] == (320, 9, 3, 3) and spec["up_blocks.1.resnets.2.conv1.weight"] == (1280, 1920, 3, 3)
    pe = sinusoidal_pe(24, 320)
    assert pe.shape == (1, 24, 320) and float(pe[0, 0, 1]) == 1.0 and abs(float(pe[0, 1, 0]) - np.sin(1.0)) < 1e-7


def test_vae_key_contract():
    from followyourclick_b200 import AutoencoderKL
    ref = {k: tuple(s) for k, s in json.load(open(os.path.join(GOLD, "vae_keys.json"))).items()}
    vae = AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                        up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=MINI_VAE["block_out_channels"],
                        layers_per_block=MINI_VAE["layers_per_block"], latent_channels=4, norm_num_groups=32)
    assert {k: tuple(v.shape) for k, v in vae.state_dict().items()} == ref
    assert tuple(vae.config.block_out_channels) == MINI_VAE["block_out_channels"]


def test_scheduler_host_tables_bit_exact_vs_reference_fixture():
    from followyourclick_b200 import DDIMScheduler
    from oracle.ref_ddim import DDIMOracle
    g = np.load(os.path.join(GOLD, "ddim.npz"))
    for name, cfg in (("v", SCHED_V), ("eps", SCHED_EPS)):
        s = DDIMScheduler(**cfg)
        assert np.array_equal(s.alphas_cumprod.numpy(), g[f"{name}_alphas_cumprod"])
        for n in (4, 25, 50):
            s.set_timesteps(n)
            assert np.array_equal(s.timesteps.numpy(), g[f"{name}_timesteps_{n}"]) and s._timesteps_host == s.timesteps.tolist()
        # coefficient scalars equal the oracle's fp32 expressions
        o = DDIMOracle(cfg)
        o.set_timesteps(25); s.set_timesteps(25)
        for t in (961, 481, 1):
            c = s.coefs(t, eta=0.3, guidance=8.0)
            a_t, prev = o.alphas_cumprod[t], t - 40
            a_prev = o.alphas_cumprod[prev] if prev >= 0 else o.final_alpha_cumprod
            assert c.sqrt_alpha_t == float(a_t ** 0.5) and c.sqrt_beta_t == float((1 - a_t) ** 0.5)
            assert c.sqrt_alpha_prev == float(a_prev ** 0.5) and c.guidance == 8.0
    assert s.init_noise_sigma == 1.0 and s.order == 1 and s.config.steps_offset == 1
    x = torch.zeros(2, 3)
    assert s.scale_model_input(x, 5) is x
    with pytest.raises(ValueError):
        DDIMScheduler(**SCHED_V).step(x, 1, x)          # set_timesteps not called (scheduling_ddim.py:291-294)
    with pytest.raises(ValueError):
        DDIMScheduler(prediction_type="bogus")


def test_geglu_interleave_roundtrip():
    from followyourclick_b200.modeling import geglu_interleave
    w, b = torch.randn(1280, 16), torch.randn(1280)
    wi, bi = geglu_interleave(w, b)
    x = torch.randn(5, 16)
    h = x @ wi.t() + bi
    t = h.view(5, 5, 2, 128)
    a, gate = t[:, :, 0].reshape(5, 640), t[:, :, 1].reshape(5, 640)
    ref_a, ref_g = (x @ w.t() + b).chunk(2, dim=-1)
    assert torch.allclose(a, ref_a, atol=1e-5) and torch.allclose(gate, ref_g, atol=1e-5)


def test_pipeline_input_validation_and_no_cpu_fallback():
    from followyourclick_b200 import AnimationPipeline, AutoencoderKL, DDIMScheduler, UNet3DConditionModel
    from tests.engine_helpers import FakeTextEncoder, FakeTokenizer, make_unet, make_vae
    unet, _ = make_unet("base", device=None)
    vae, _ = make_vae(device=None)
    pipe = AnimationPipeline(vae=vae, text_encoder=FakeTextEncoder(torch.zeros(2, 77, 768)), tokenizer=FakeTokenizer(),
                             unet=unet, scheduler=DDIMScheduler(**SCHED_V))
    assert pipe.vae_scale_factor == 8
    with pytest.raises(ValueError):
        pipe(prompt="p", video_length=4, height=60, width=64)                       # pipeline_animation.py:436-437
    with pytest.raises(ValueError):
        pipe(prompt=3, video_length=4, height=64, width=64)
    with pytest.raises(ValueError):
        pipe.prepare_latents(1, 4, 4, 64, 64, torch.float32, "cpu", None, latents=torch.zeros(1, 4, 4, 9, 8))   # :517-518
    # the engine never silently computes on the CPU
    with pytest.raises(RuntimeError):
        unet.forward_nfhwc(torch.zeros(2, 4, 16, 16, 9), 1, torch.zeros(2, 77, 768))
    with pytest.raises(RuntimeError):
        vae.decode_nhwc(torch.zeros(1, 8, 8, 4))
    with pytest.raises(RuntimeError):
        vae.encode(torch.zeros(1, 3, 64, 64))


def test_synth_weights_are_deterministic_and_nonzero():
    from followyourclick_b200.synth import synth_state_dict
    shapes = {"a.to_q.weight": (8, 4), "b.proj_out.weight": (4, 4), "b.proj_out.bias": (4,), "n.norm.weight": (6,),
              "m.pos_encoder.pe": (1, 2, 3)}
    s1, s2 = synth_state_dict(shapes), synth_state_dict(shapes)
    assert set(s1) == set(shapes) - {"m.pos_encoder.pe"}
    assert all(torch.equal(s1[k], s2[k]) for k in s1) and float(s1["b.proj_out.weight"].abs().sum()) > 0
    assert abs(float(s1["n.norm.weight"].mean()) - 1.0) < 0.3


def test_upsample_phase_weights_restates_upsample_plus_conv():
    """modeling.upsample_phase_weights: nearest-x2 + padded 3x3 conv == four 2x2-tap convs on the low-res image (fp64 check of the
    restatement the tcgen05 upsampler path runs; the kernel-level parity is tests/test_kernels_gpu.py)."""
    import torch.nn.functional as F
    from followyourclick_b200.modeling import upsample_phase_weights
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 6, 7, generator=g, dtype=torch.float64)
    w = torch.randn(4, 5, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    wp = upsample_phase_weights(w.float()).double()
    assert wp.shape == (4, 4, 2, 2, 5)
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for py in range(2):
        for px in range(2):
            acc = 0
            for a in range(2):
                for b in range(2):
                    dy, dx = a - 1 + py, b - 1 + px
                    acc = acc + torch.einsum("nchw,oc->nohw", xp[:, :, 1 + dy:7 + dy, 1 + dx:8 + dx], wp[2 * py + px, :, a, b, :])
            out[:, :, py::2, px::2] = acc
    assert float((out - ref).abs().max()) < 1e-5


def test_ip_adapter_weight_surgery_pairs_keys_by_order():
    """MyIPAdapter.load_ip_adapter (ip_adapter/my_ip_adapter.py:85-125): projector weights into unet.image_proj_model, the adapter
    file's to_k_ip / to_v_ip tensors into the UNet's *_ip* tensors in key order."""
    from followyourclick_b200 import MyIPAdapter, UNet3DConditionModel
    unet = UNet3DConditionModel(**mini_unet_ref_kwargs("ip"))
    ad = MyIPAdapter(unet, image_encoder=object(), device="cpu", num_tokens=4, clip_embeddings_dim=1024)
    ip_keys = [k for k in unet.state_dict() if "_ip" in k]
    assert len(ip_keys) == 2 * 10                                  # to_k_ip + to_v_ip of the 10 transformer blocks of the mini model
    g = torch.Generator().manual_seed(0)
    sd = {"image_proj": {k: torch.randn(v.shape, generator=g) for k, v in ad.image_proj_model.state_dict().items()},
          "ip_adapter": {f"{i}.w": torch.randn(unet.state_dict()[k].shape, generator=g) for i, k in enumerate(ip_keys)}}
    missing, unexpected = ad.load_ip_adapter(unet, use_unet_image_proj_model=True, state_dict=sd)
    assert not unexpected
    for i, k in enumerate(ip_keys):
        assert torch.equal(unet.state_dict()[k], sd["ip_adapter"][f"{i}.w"])
    assert torch.equal(unet.image_proj_model.state_dict()["proj.weight"], sd["image_proj"]["proj.weight"])
