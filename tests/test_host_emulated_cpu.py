"""CPU: the HOST logic of the drop-in classes, with every kernel launch replaced by tests/ops_emulator.py (torch restatements of
the documented kernel contracts).  The product has no CPU path; these tests exist so that weight packing, layout plumbing, the
ClipContext hoisting, the Resampler and the pipeline loop are checked against the reference fixtures in the GPU-less build
container too.  The real kernels are checked by the ``-m gpu`` suites through the C ABI.

Tolerances are the engine's own (tests/test_engine_gpu.py): fp32 rel-L2 <= 1e-4, bf16 rel-L2 <= 3e-2, video PSNR >= 30 dB.
"""
import pytest
import torch

from tests import ops_emulator
from tests.cfgs import MINI_UNET_VARIANTS

DTYPES = [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)]


@pytest.fixture(autouse=True)
def _emulated(monkeypatch):
    ops_emulator.install(monkeypatch)
    torch.set_num_threads(8)
    yield


@pytest.mark.parametrize("variant", MINI_UNET_VARIANTS)
@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_unet_host_logic_vs_reference_golden(variant, dtype, tol):
    from tests.engine_helpers import run_unet_case
    s = run_unet_case(variant, dtype, device="cpu")
    assert s["finite"] and s["rel_l2"] < tol, s


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_vae_host_logic_vs_reference_golden(dtype, tol):
    from tests.engine_helpers import run_vae_case, run_vae_encode_case
    s = run_vae_case(dtype, device="cpu")
    assert s["finite"] and s["rel_l2"] < tol, s
    m, smp, dist = run_vae_encode_case(dtype, device="cpu")
    assert m["rel_l2"] < tol and smp["rel_l2"] < tol, (m, smp)


def test_pipeline_host_logic_vs_reference_golden():
    from tests.engine_helpers import run_pipeline_case
    r = run_pipeline_case(torch.float32, steps=3, against="golden", device="cpu")
    assert r["finite"] and r["shape"] == (1, 3, 4, 64, 64) and r["video_maxabs"] < 2e-3, r
    r = run_pipeline_case(torch.bfloat16, steps=3, against="golden", device="cpu")
    assert r["finite"] and r["psnr"] > 30.0, r


def test_upsampler_phase_path_and_padded_heads_are_the_ones_exercised(monkeypatch):
    """bf16 mode must route the upsamplers through w_phases and the 3- / 4-channel heads through 16 padded channels; switching
    either off (FYC_UP2_PHASES=0 / FYC_TC_HEAD=0 equivalents) changes the result only by bf16 rounding."""
    from followyourclick_b200 import ops
    from tests.engine_helpers import run_unet_case, run_vae_case
    calls = {"phases": 0, "head16": 0}
    conv = ops.conv3x3

    def spy(x, w, *a, **kw):
        calls["phases"] += kw.get("w_phases") is not None
        calls["head16"] += (w.shape[0] == 16 and kw.get("w_phases") is None)
        return conv(x, w, *a, **kw)
    monkeypatch.setattr(ops, "conv3x3", spy)
    a = run_unet_case("base", torch.bfloat16, device="cpu")
    v = run_vae_case(torch.bfloat16, device="cpu")
    assert calls["phases"] == 3 + 3 and calls["head16"] == 2, calls      # 3 upsamplers each; conv_out of each model
    monkeypatch.setattr(ops, "use_up2_phases", False)
    monkeypatch.setattr(ops, "use_tc_head", False)
    b = run_unet_case("base", torch.bfloat16, device="cpu")
    v2 = run_vae_case(torch.bfloat16, device="cpu")
    assert abs(a["rel_l2"] - b["rel_l2"]) < 1e-2 and abs(v["rel_l2"] - v2["rel_l2"]) < 1e-2, (a, b, v, v2)


def test_resampler_host_logic_vs_reference_golden():
    from followyourclick_b200 import Resampler
    from tests.cfgs import MINI_RESAMPLER
    from tests.engine_helpers import golden, load_synth, stats
    m = Resampler(**MINI_RESAMPLER)
    load_synth(m)
    g = golden("resampler.npz")
    s = stats(m(torch.from_numpy(g["x"])), torch.from_numpy(g["out"]))
    assert s["finite"] and s["rel_l2"] < 1e-5, s


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_context_hoisting_and_ip_plus_host_logic(dtype, tol):
    """prepare_context once == rebuilding per forward (bit for bit); UNet + Resampler projector (MyIPAdapterPlus) vs the oracle."""
    from followyourclick_b200 import Resampler, UNet3DConditionModel, ops
    from oracle import ref_resampler, ref_unet
    from tests.cfgs import MINI_RESAMPLER, RESAMPLER_TOKENS, mini_unet_oracle_cfg, mini_unet_ref_kwargs, unet_inputs
    from tests.engine_helpers import load_synth, stats
    kw = dict(mini_unet_ref_kwargs("ip"), num_tokens=MINI_RESAMPLER["num_queries"])
    unet = UNet3DConditionModel(**kw)
    sd = load_synth(unet)
    unet.to(dtype)
    rs = Resampler(**MINI_RESAMPLER)
    rsd = load_synth(rs)
    unet.image_proj_model = rs
    inp = unet_inputs("ip")
    clip = torch.randn(2, RESAMPLER_TOKENS, MINI_RESAMPLER["embedding_dim"], generator=torch.Generator().manual_seed(5))
    x = ops.ncfhw_to_nfhwc(inp["sample"].contiguous(), dtype)
    args = dict(fps_tensor=inp["fps"], flow_control=inp["flow"], use_fps_condition=True, use_ip_cross_attention=True)
    a = unet.forward_nfhwc(x, inp["timestep"], inp["ctx"], reference_images_clip_feat=clip, **args)
    ctx = unet.prepare_context(inp["ctx"], clip, True)
    b = unet.forward_nfhwc(x, inp["timestep"], None, context=ctx, **args)
    assert torch.equal(a, b)
    # every block's context projections are hoisted: packed for the tcgen05 cross-attention (head dims 40 / 80, tensor-core mode) or plain [K | V]
    assert len(ctx.kv) + len(ctx.kx) == len(ctx.kvi) + len(ctx.kxi) == len(unet._transformer_prefixes()) == 10
    assert (len(ctx.kx) == 6) == (dtype == torch.bfloat16)
    tokens = ref_resampler.resampler_forward(rsd, MINI_RESAMPLER, clip)
    ocfg = dict(mini_unet_oracle_cfg("ip"), num_tokens=MINI_RESAMPLER["num_queries"])
    ref = ref_unet.unet3d_forward(sd, ocfg, inp["sample"], inp["timestep"], torch.cat([inp["ctx"], tokens], dim=1),
                                  fps_tensor=inp["fps"], flow_control=inp["flow"], reference_images_clip_feat=None)
    s = stats(ops.nfhwc_to_ncfhw(a), ref)
    assert s["finite"] and s["rel_l2"] < tol, s


def test_pipeline_hoisted_equals_per_step_host_logic():
    from followyourclick_b200 import AnimationPipeline
    from tests.engine_helpers import make_pipeline, pipeline_call
    vids = []
    for hoist in (True, False):
        pipe, ci, _, _ = make_pipeline(torch.bfloat16, device="cpu")
        pipe.use_cuda_graph = False
        AnimationPipeline.hoist_context = hoist
        try:
            vids.append(pipeline_call(pipe, ci, 4, 8, 8, 2, 8.0))
        finally:
            AnimationPipeline.hoist_context = True
    assert torch.equal(vids[0], vids[1])


def test_video_scale_branch_host_logic_vs_reference_golden():
    """per-frame guidance (video_scale > 0) through the product pipeline (eager loop) vs the reference's frames."""
    from tests.engine_helpers import run_video_scale_case
    r = run_video_scale_case(torch.float32, device="cpu")
    assert r["finite"] and r["shape"] == (1, 3, 4, 64, 64) and r["video_maxabs"] < 2e-3, r
    r = run_video_scale_case(torch.bfloat16, device="cpu")
    assert r["finite"] and r["psnr"] > 30.0, r


@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_unet2d_host_logic_vs_reference_golden(dtype, tol):
    from tests.engine_helpers import run_unet2d_case
    s = run_unet2d_case(dtype, device="cpu")
    assert s["finite"] and s["rel_l2"] < tol, s


def test_unet_ragged_shape_host_logic():
    from tests.engine_helpers import run_unet_ragged_case
    s = run_unet_ragged_case(torch.bfloat16, device="cpu")
    assert s["finite"] and s["rel_l2"] < 3e-2, s


@pytest.mark.parametrize("variant", MINI_UNET_VARIANTS)
@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_shared_cfg_prefix_equals_duplicated_batch(variant, dtype, tol):
    """forward_nfhwc(cfg_dup=2) on ONE copy of the input == the reference's duplicated batch (both halves carry identical values up to
    the first cross-attention, pipeline_animation.py:709), and therefore matches the reference fixture computed on [x, x]."""
    from followyourclick_b200 import ops
    from tests.engine_helpers import golden, make_unet, stats, unet_forward_kwargs
    from tests.cfgs import unet_inputs
    unet, _ = make_unet(variant, dtype, "cpu")
    inp = unet_inputs(variant)
    one = inp["sample"][:1]                                   # the fixture's batch is two DIFFERENT samples: use a CFG-style pair of sample 0
    kw = unet_forward_kwargs(variant, inp, "cpu")
    nf = dict(fps_tensor=kw["fps_tensor"], flow_control=kw["flow_control"], reference_images_clip_feat=kw["reference_images_clip_feat"],
              camera_movement_type_tensor=kw["camera_movement_type_tensor"], use_ip_cross_attention=kw["use_ip_cross_attention"],
              use_camera_motion_condition=kw["use_camera_motion_condition"], use_fps_condition=kw["use_fps_condition"])
    x1 = ops.ncfhw_to_nfhwc(one.contiguous(), dtype)
    full = unet.forward_nfhwc(torch.cat([x1, x1]), inp["timestep"], kw["encoder_hidden_states"], **nf)
    shared = unet.forward_nfhwc(x1, inp["timestep"], kw["encoder_hidden_states"], cfg_dup=2, **nf)
    assert shared.shape == full.shape
    s = stats(shared, full)
    assert s["finite"] and s["rel_l2"] < (1e-6 if dtype == torch.float32 else 2e-3), s
    assert not torch.equal(full[0], full[1])                  # the two halves do differ after the cross-attention (different text rows)


@pytest.mark.parametrize("share", [True, False])
def test_pipeline_with_shared_cfg_prefix_vs_reference_golden(share):
    """both settings of AnimationPipeline.share_cfg_prefix (default on; FYC_SHARED_PREFIX=0 duplicates the CFG batch like the reference)"""
    from followyourclick_b200 import AnimationPipeline
    from tests.engine_helpers import run_pipeline_case
    _old_share = AnimationPipeline.share_cfg_prefix
    AnimationPipeline.share_cfg_prefix = share
    try:
        r = run_pipeline_case(torch.float32, steps=3, against="golden", device="cpu")
    finally:
        AnimationPipeline.share_cfg_prefix = _old_share
    assert r["finite"] and r["video_maxabs"] < 2e-3, r


@pytest.mark.parametrize("variant", ["ip", "cam"])
def test_pipeline_ip_and_camera_variants_host_logic_vs_reference_golden(variant):
    """configs[2] / [4] plumbing: [uncond, cond] image features, camera-motion embedding, epsilon prediction, 4-channel input."""
    from tests.engine_helpers import run_pipeline_variant_case
    r = run_pipeline_variant_case(variant, torch.float32, device="cpu")
    assert r["finite"] and r["shape"] == (1, 3, 4, 64, 64) and r["video_maxabs"] < 2e-3, r


@pytest.fixture
def graph_branch(monkeypatch):
    """Walk AnimationPipeline's CUDA-graph branch on CPU: _GraphedUNetStep in capture-free mode (replay() re-executes the forward)."""
    from followyourclick_b200.pipeline_animation import _GraphedUNetStep
    monkeypatch.setattr(_GraphedUNetStep, "capture", False)


def test_graph_branch_bookkeeping_vs_reference_golden(graph_branch):
    """Static input buffers, the graph cache, ClipContext refresh on a second clip, the video_scale step's VIEW of the CFG input, the
    IP / camera variants and the shared CFG prefix - the host code of the graph branch, against the reference fixtures."""
    from followyourclick_b200 import AnimationPipeline
    from tests.engine_helpers import (make_pipeline, pipeline_call, run_pipeline_case, run_pipeline_variant_case, run_video_scale_case,
                                      golden)
    r = run_pipeline_case(torch.float32, steps=3, against="golden", device="cpu")
    assert r["video_maxabs"] < 2e-3, r
    r = run_video_scale_case(torch.float32, device="cpu", graph=True)
    assert r["video_maxabs"] < 2e-3, r
    for v in ("ip", "cam"):
        r = run_pipeline_variant_case(v, torch.float32, device="cpu", graph=True)
        assert r["video_maxabs"] < 2e-3, (v, r)
    _old_share = AnimationPipeline.share_cfg_prefix
    AnimationPipeline.share_cfg_prefix = True
    try:
        r = run_pipeline_case(torch.float32, steps=3, against="golden", device="cpu")
        assert r["video_maxabs"] < 2e-3, r
        r = run_video_scale_case(torch.float32, device="cpu", graph=True)
        assert r["video_maxabs"] < 2e-3, r
    finally:
        AnimationPipeline.share_cfg_prefix = _old_share
    # a second clip through the SAME pipeline object reuses the cached step and must refresh its static buffers / context
    pipe, ci, _, _ = make_pipeline(torch.float32, device="cpu")
    pipe.use_cuda_graph = True
    a = pipeline_call(pipe, ci, 4, 8, 8, 3, 8.0)
    ci2 = dict(ci, text_embeddings=ci["text_embeddings"] * 0.5)
    pipe.text_encoder.emb = ci2["text_embeddings"]
    b = pipeline_call(pipe, ci2, 4, 8, 8, 3, 8.0)
    pipe.text_encoder.emb = ci["text_embeddings"]
    c = pipeline_call(pipe, ci, 4, 8, 8, 3, 8.0)
    assert len(pipe._graph_cache) == 1 and torch.equal(a, c) and not torch.equal(a, b)
    assert float((a - torch.from_numpy(golden("pipeline.npz")["video"])).abs().max()) < 2e-3


def test_dry_run_of_the_full_size_property_tests(monkeypatch):
    """tests/test_zz_late_gpu.py's full-size identities, executed here at reduced sizes on the emulated kernels: checks that the
    identities themselves (and the test code) are right; on a GPU box the same functions run at cfg2 size on the real kernels."""
    import tests.test_zz_late_gpu as late
    monkeypatch.setattr(late, "DEV", "cpu")
    monkeypatch.setattr(late, "SMALL", True)
    late.test_full_size_kernel_identities()
    late.test_full_size_unet_and_pipeline_properties()


def test_shared_cfg_prefix_with_two_clips():
    """b = 2 clips: context rows are [uncond clip 0, uncond clip 1, cond clip 0, cond clip 1]; the prefix runs on the 2 clips once."""
    from followyourclick_b200 import ops
    from tests.cfgs import unet_inputs
    from tests.engine_helpers import make_unet, stats
    unet, _ = make_unet("base", torch.float32, "cpu")
    inp = unet_inputs("base", b=2)
    x2 = ops.ncfhw_to_nfhwc(inp["sample"].contiguous(), torch.float32)
    ctx4 = torch.randn(4, 77, 768, generator=torch.Generator().manual_seed(8))
    kw = dict(fps_tensor=torch.tensor([2] * 4), flow_control=torch.tensor([4] * 4), use_fps_condition=True)
    full = unet.forward_nfhwc(torch.cat([x2, x2]), inp["timestep"], ctx4, **kw)
    shared = unet.forward_nfhwc(x2, inp["timestep"], ctx4, cfg_dup=2, **kw)
    s = stats(shared, full)
    assert shared.shape == full.shape == (4, 4, 16, 16, 4) and s["rel_l2"] < 1e-6, s


@pytest.mark.parametrize("case", ["tok_t4", "img_t16", "res_t4"])
def test_ip_attn_processor_host_logic_vs_reference_processor(case):
    """host side of IPAttnProcessor (token split [:, :-T] / [:, -T:], 4-D <-> token reshapes, residual_connection, the fused
    second-context call) against the unmodified reference processor's fixture, kernels emulated"""
    from tests.engine_helpers import run_ip_attn_processor_case
    s = run_ip_attn_processor_case(case, torch.float32, device="cpu")
    assert s["finite"] and s["rel_l2"] < 1e-5, s


def test_smoke_entry_logic_emulated():
    """`__graft_entry__.smoke()` minus the device: the same three helper calls with the same thresholds, kernels emulated (its oracle leg -
    `run_pipeline_case(..., steps=1, against="oracle")` - is not what the `-m gpu` engine tests run, so it is kept alive here)."""
    from tests.engine_helpers import run_pipeline_case, run_unet_case
    u16 = run_unet_case("base", torch.bfloat16, device="cpu")
    assert u16["finite"] and u16["rel_l2"] < 3e-2, u16
    r16 = run_pipeline_case(dtype=torch.bfloat16, steps=1, device="cpu")
    assert r16["video_maxabs"] < 0.1, r16
    r = run_pipeline_case(dtype=torch.float32, steps=1, device="cpu")
    assert r["video_maxabs"] < 2e-3, r
