"""CPU: the oracle (oracle/) re-checked against the reference-generated golden fixtures on every run."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_ddim, ref_pipeline, ref_unet, ref_vae
from tests.cfgs import MINI_UNET_VARIANTS, MINI_VAE, SCHED_EPS, SCHED_V, mini_unet_oracle_cfg, unet_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _synth(keys):
    from followyourclick_b200.synth import synth_state_dict
    from followyourclick_b200.unet import sinusoidal_pe
    sd = synth_state_dict({k: tuple(s) for k, s in keys.items()})
    for k, s in keys.items():
        if k.endswith(".pos_encoder.pe"):
            sd[k] = sinusoidal_pe(s[1], s[2])
    return sd


def test_ddim_known_answers_from_survey_appendix_d():
    o = ref_ddim.DDIMOracle(SCHED_V)
    assert o.set_timesteps(4).tolist() == [751, 501, 251, 1]
    assert o.set_timesteps(25).tolist()[:2] == [961, 921] and o.set_timesteps(50).tolist()[:2] == [981, 961]
    ac = o.alphas_cumprod
    np.testing.assert_allclose(ac[[0, 1, 2]].numpy(), [0.9991499782, 0.9982538819, 0.9973470569], rtol=1e-6)
    np.testing.assert_allclose(float(ac[481]), 0.16172254, rtol=1e-5)
    np.testing.assert_allclose(float(ac[961]), 1.0946615e-4, rtol=1e-4)
    assert float(ac[999]) == 0.0
    e = ref_ddim.DDIMOracle(SCHED_EPS).alphas_cumprod
    np.testing.assert_allclose([float(e[1]), float(e[961]), float(e[999])], [0.9982895255, 0.0024783323, 0.0015789628], rtol=1e-5)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 4, 16, 64, 64, generator=g)
    v = torch.randn(1, 4, 16, 64, 64, generator=g)
    o.set_timesteps(25)
    r = o.step(v, 961, x)
    np.testing.assert_allclose(float(r.double().sum()), 1200.6520849, rtol=1e-6)
    np.testing.assert_allclose(r.flatten()[:3].numpy(), [-0.1251385212, -0.4922964275, 0.1633480191], rtol=1e-6)
    r = o.step(v, 1, x)
    np.testing.assert_allclose(float(r.double().sum()), 1195.7192700, rtol=1e-6)


def test_ddim_oracle_bit_exact_vs_reference_fixture():
    g = np.load(os.path.join(GOLD, "ddim.npz"))
    for name, cfg in (("v", SCHED_V), ("eps", SCHED_EPS)):
        o = ref_ddim.DDIMOracle(cfg)
        assert np.array_equal(o.alphas_cumprod.numpy(), g[f"{name}_alphas_cumprod"])
        x, v = torch.from_numpy(g[f"{name}_x"]), torch.from_numpy(g[f"{name}_v"])
        for n in (4, 25, 50):
            ts = o.set_timesteps(n)
            assert np.array_equal(ts.numpy(), g[f"{name}_timesteps_{n}"])
            for t in (int(ts[0]), int(ts[len(ts) // 2]), int(ts[-1])):
                assert np.array_equal(o.step(v, t, x).numpy(), g[f"{name}_step_{n}_{t}"])


@pytest.mark.parametrize("variant", MINI_UNET_VARIANTS)
def test_unet_oracle_vs_reference_fixture(variant):
    keys = json.load(open(os.path.join(GOLD, "unet_keys.json")))[variant]
    sd = _synth(keys)
    inp = unet_inputs(variant)
    out = ref_unet.unet3d_forward(sd, mini_unet_oracle_cfg(variant), inp["sample"], inp["timestep"], inp["ctx"],
                                  fps_tensor=inp.get("fps"), flow_control=inp.get("flow"),
                                  reference_images_clip_feat=inp.get("clip"), camera_movement_type_tensor=inp.get("camera"))
    ref = torch.from_numpy(np.load(os.path.join(GOLD, f"unet_{variant}.npz"))["out"])
    assert float((out - ref).abs().max()) < 5e-4 * max(1.0, float(ref.abs().max()))


def test_vae_encode_oracle_vs_reference_fixture():
    """the encoder restatement against the unmodified reference's moments and reparameterised sample."""
    vsd = _synth(json.load(open(os.path.join(GOLD, "vae_keys.json"))))
    g = np.load(os.path.join(GOLD, "vae_encode.npz"))
    m = ref_vae.vae_encode_moments(vsd, MINI_VAE, torch.from_numpy(g["x"]))
    assert m.shape == (2, 8, 8, 8) and float((m - torch.from_numpy(g["moments"])).abs().max()) < 1e-4
    s = ref_vae.gaussian_sample(m, torch.from_numpy(g["noise"]))
    assert float((s - torch.from_numpy(g["sample"])).abs().max()) < 1e-4


def test_vae_and_pipeline_oracle_vs_reference_fixture():
    from followyourclick_b200.synth import synth_clip_inputs
    vkeys = json.load(open(os.path.join(GOLD, "vae_keys.json")))
    vsd = _synth(vkeys)
    g = np.load(os.path.join(GOLD, "vae.npz"))
    out = ref_vae.vae_decode(vsd, MINI_VAE, torch.from_numpy(g["z"]))
    assert float((out - torch.from_numpy(g["out"])).abs().max()) < 1e-3
    usd = _synth(json.load(open(os.path.join(GOLD, "unet_keys.json")))["base"])
    ci = synth_clip_inputs(1, 4, 8, 8)
    lat = ref_pipeline.denoise(usd, mini_unet_oracle_cfg("base"), SCHED_V, ci["latents"], ci["text_embeddings"], 3, 8.0,
                               first_image_latents=ci["first_image_latents"], first_images_mask=ci["first_images_mask"],
                               fps_tensor=torch.tensor([2]), flow_control=torch.tensor([4]))
    p = np.load(os.path.join(GOLD, "pipeline.npz"))
    assert float((lat - torch.from_numpy(p["final_latents"])).abs().max()) < 1e-3
    video = ref_vae.decode_latents(vsd, MINI_VAE, lat)
    assert float((video - torch.from_numpy(p["video"])).abs().max()) < 2e-3


def test_resampler_oracle_vs_reference_fixture():
    """oracle restatement of the Perceiver Resampler against the unmodified reference's output."""
    from oracle import ref_resampler
    from followyourclick_b200.synth import synth_state_dict
    from tests.cfgs import MINI_RESAMPLER, RESAMPLER_TOKENS
    g = np.load(os.path.join(GOLD, "resampler.npz"))
    sd = synth_state_dict(ref_resampler.resampler_param_shapes(MINI_RESAMPLER))
    x = torch.from_numpy(g["x"])
    assert x.shape == (2, RESAMPLER_TOKENS, MINI_RESAMPLER["embedding_dim"])
    out = ref_resampler.resampler_forward(sd, MINI_RESAMPLER, x)
    assert out.shape == (2, MINI_RESAMPLER["num_queries"], MINI_RESAMPLER["output_dim"])
    assert float((out - torch.from_numpy(g["out"])).abs().max()) < 2e-5
    pins = json.load(open(os.path.join(GOLD, "pins.json")))["oracle_vs_reference_maxabs"]
    assert pins["resampler"] < 2e-5


def test_pipeline_video_scale_oracle_vs_reference_fixture():
    """oracle's per-frame guidance branch (pipeline_animation.py:738-761) against the reference's frames."""
    from followyourclick_b200.synth import synth_clip_inputs
    from tests.cfgs import MINI_VAE as VCFG
    g = np.load(os.path.join(GOLD, "pipeline_video_scale.npz"))
    keys = json.load(open(os.path.join(GOLD, "unet_keys.json")))["base"]
    vkeys = json.load(open(os.path.join(GOLD, "vae_keys.json")))
    usd, vsd = _synth(keys), _synth(vkeys)
    ci = synth_clip_inputs(1, 4, 8, 8)
    lat = ref_pipeline.denoise(usd, mini_unet_oracle_cfg("base"), SCHED_V, ci["latents"], ci["text_embeddings"], int(g["steps"]), 8.0,
                               first_image_latents=ci["first_image_latents"], first_images_mask=ci["first_images_mask"],
                               fps_tensor=torch.tensor([2]), flow_control=torch.tensor([4]), video_scale=float(g["video_scale"]))
    assert float((lat - torch.from_numpy(g["final_latents"])).abs().max()) < 1e-4
    video = ref_vae.decode_latents(vsd, VCFG, lat)
    assert float((video - torch.from_numpy(g["video"])).abs().max()) < 2e-4


def test_unet2d_oracle_vs_reference_fixture():
    """oracle (3-D restatement, no motion modules, one frame) vs the reference's 2-D UNet output."""
    from tests.cfgs import mini_unet2d_oracle_cfg
    g = np.load(os.path.join(GOLD, "unet2d.npz"))
    sd = _synth(json.load(open(os.path.join(GOLD, "unet2d_keys.json"))))
    out = ref_unet.unet3d_forward(sd, mini_unet2d_oracle_cfg(), torch.from_numpy(g["x"]).unsqueeze(2), torch.tensor(501),
                                  torch.from_numpy(g["ctx"])).squeeze(2)
    assert float((out - torch.from_numpy(g["out"])).abs().max()) < 1e-4


def test_video_grid_oracle_pinned_against_torchvision():
    """the oracle's make_grid == torchvision.utils.make_grid (what save_videos_grid calls) for 1, 3 and 7 clips."""
    torchvision = pytest.importorskip("torchvision")
    from oracle import ref_util
    g = torch.Generator().manual_seed(4)
    for b, nrow in ((1, 6), (3, 6), (7, 4), (4, 4)):
        x = torch.rand(b, 3, 5, 6, generator=g)
        assert torch.equal(ref_util.make_grid(x, nrow), torchvision.utils.make_grid(x, nrow=nrow))
    v = torch.rand(3, 3, 2, 5, 6, generator=g)
    fr = ref_util.video_frames_uint8(v, n_rows=2)
    assert len(fr) == 2 and fr[0].shape == (2 * 7 + 2, 2 * 8 + 2, 3) and fr[0].dtype == np.uint8


@pytest.mark.parametrize("variant", ["ip", "cam"])
def test_pipeline_variant_oracle_vs_reference_fixture(variant):
    """configs[2] / [4] plumbing (IP-Adapter image condition; camera-LoRA model with epsilon prediction) - oracle vs reference frames."""
    from tests.cfgs import CLIP_DIM, pipeline_variant_inputs
    g = np.load(os.path.join(GOLD, f"pipeline_{variant}.npz"))
    keys = dict(json.load(open(os.path.join(GOLD, "unet_keys.json")))[variant])
    usd = _synth(keys)
    ci, _, okw, sched_cfg, steps, gs = pipeline_variant_inputs(variant)
    lat = ref_pipeline.denoise(usd, mini_unet_oracle_cfg(variant), sched_cfg, ci["latents"], ci["text_embeddings"], steps, gs,
                               image_clip_feat=ci["image_clip_feat"], uncond_image_clip_feat=ci["uncond_image_clip_feat"], **okw)
    assert float((lat - torch.from_numpy(g["final_latents"])).abs().max()) < 1e-4


def test_unet_oracle_vs_reference_fixture_ragged_shape():
    """Oracle pinned on a non-square, non-power-of-two grid with an odd frame count and batch 1 (reference run on 1 x 3 x 24 x 40)."""
    g = np.load(os.path.join(GOLD, "unet_base_ragged.npz"))
    sd = _synth(json.load(open(os.path.join(GOLD, "unet_keys.json")))["base"])
    inp = unet_inputs("base", b=1, f=3, h=24, w=40, seed=23)
    out = ref_unet.unet3d_forward(sd, mini_unet_oracle_cfg("base"), inp["sample"], inp["timestep"], inp["ctx"],
                                  fps_tensor=torch.tensor([2]), flow_control=torch.tensor([4]))
    assert float((out - torch.from_numpy(g["out"])).abs().max()) < 1e-4
