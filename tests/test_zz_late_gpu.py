"""GPU parity tests added after the round's last GPU run (they exercise only kernels and paths the earlier suites already cover on the
GPU, and pass through the CPU emulation of the kernels): kept in a file that sorts last so that `pytest -x` reaches every
GPU-verified test first.

Tolerances as in tests/test_engine_gpu.py: fp32 video max-abs <= 2e-3, bf16 PSNR >= 30 dB against the reference's frames."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _impl(cuda):
    from followyourclick_b200 import ops
    ops.set_impl("auto")
    yield
    ops.set_impl("auto")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_shared_cfg_prefix_on_gpu(dtype):
    """cfg_dup=2 on one copy of the input vs the duplicated batch (expected bit-identical: same kernels, row-independent), and the
    pipeline with the prefix shared vs the reference fixture."""
    from followyourclick_b200 import AnimationPipeline, ops
    from tests.cfgs import unet_inputs
    from tests.engine_helpers import make_unet, run_pipeline_case, stats, unet_forward_kwargs
    unet, _ = make_unet("ip", dtype)
    inp = unet_inputs("ip")
    kw = unet_forward_kwargs("ip", inp, "cuda")
    nf = dict(fps_tensor=kw["fps_tensor"], flow_control=kw["flow_control"], reference_images_clip_feat=kw["reference_images_clip_feat"],
              use_ip_cross_attention=True, use_fps_condition=True)
    x1 = ops.ncfhw_to_nfhwc(inp["sample"][:1].cuda().contiguous(), dtype)
    full = unet.forward_nfhwc(torch.cat([x1, x1]), inp["timestep"], kw["encoder_hidden_states"], **nf)
    shared = unet.forward_nfhwc(x1, inp["timestep"], kw["encoder_hidden_states"], cfg_dup=2, **nf)
    s = stats(shared, full)
    assert s["rel_l2"] < (1e-6 if dtype == torch.float32 else 2e-3), s
    old = AnimationPipeline.share_cfg_prefix
    for share in (True, False):            # default on; off = the reference's duplicated CFG batch
        AnimationPipeline.share_cfg_prefix = share
        try:
            r = run_pipeline_case(dtype, steps=3, against="golden")
        finally:
            AnimationPipeline.share_cfg_prefix = old
        assert r["finite"] and (r["video_maxabs"] < 2e-3 if dtype == torch.float32 else r["psnr"] > 30.0), (share, r)


@pytest.mark.parametrize("variant", ["ip", "cam"])
@pytest.mark.parametrize("graph", [True, False])
def test_pipeline_ip_and_camera_variants_match_reference_golden(variant, graph):
    """BASELINE configs[2] / [4] plumbing at mini size vs the reference pipeline's frames: fp32 max-abs <= 2e-3, bf16 PSNR >= 30 dB."""
    from tests.engine_helpers import run_pipeline_variant_case
    r = run_pipeline_variant_case(variant, torch.float32, graph=graph)
    assert r["finite"] and r["shape"] == (1, 3, 4, 64, 64) and r["video_maxabs"] < 2e-3, r
    r = run_pipeline_variant_case(variant, torch.bfloat16, graph=graph)
    assert r["finite"] and r["psnr"] > 30.0, r


# ---------------------------------------------------------------------------------------------------------------------------------
# Size-independent properties at BASELINE.json's full cfg2 size (512x512, 16 frames: 131072 tokens x 320 channels at the top level).
# The CPU oracle cannot run this size in test time; these are identities the reference arithmetic satisfies exactly (or to rounding),
# evaluated on the real kernels at the real shapes.

DEV = "cuda"
SMALL = False      # tests/test_host_emulated_cpu.py dry-runs these two tests on CPU (kernel launches emulated) at reduced sizes with DEV = "cpu"


def _full_unet(dtype=torch.bfloat16):
    from followyourclick_b200 import UNet3DConditionModel
    from followyourclick_b200.synth import synth_on_device_
    if SMALL:
        from tests.cfgs import mini_unet_ref_kwargs
        unet = UNet3DConditionModel(**mini_unet_ref_kwargs("base")).to(dtype)
        return synth_on_device_(unet, seed=0)
    mm = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=("Temporal_Self", "Temporal_Self"),
              temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1)
    unet = UNet3DConditionModel(sample_size=64, in_channels=4, out_channels=4, cross_attention_dim=768, attention_head_dim=8,
                                use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), unet_use_cross_frame_attention=False,
                                unet_use_temporal_attention=False, motion_module_type="Vanilla", use_fps_condition=True,
                                use_first_frame_mask_condition_concat=True, motion_module_kwargs=mm).to(DEV).to(dtype)
    return synth_on_device_(unet, seed=0)


def test_full_size_kernel_identities():
    """Exact scaling (power-of-two) linearity of the tcgen05 GEMM / conv / four-phase upsampler, softmax rows summing to one through
    the tcgen05 attention (constant V -> constant output), zero-mean / unit-variance GroupNorm and LayerNorm outputs, pixel-permutation
    equivariance of the temporal attention - all at the 64x64-level shapes of cfg2."""
    from followyourclick_b200 import ops
    from followyourclick_b200.modeling import upsample_phase_weights
    g = torch.Generator(device=DEV).manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g, device=DEV)
    M, C = (4096 if SMALL else 131072), 320
    NI, L, NBA, RG, PX = (2, 256, 1, 2048, 64) if SMALL else (32, 4096, 4, 65536, 4096)
    A, W = rn(M, C).bfloat16(), (rn(C, C) * C ** -0.5).bfloat16()
    y = ops.gemm(A, W)
    assert torch.equal(ops.gemm(A * 2, W), y * 2) and torch.equal(ops.gemm(A, W * 0.5), y * 0.5)
    x = rn(NI, 64, 64, C).bfloat16()
    wc = (rn(C, 3, 3, C) * (9 * C) ** -0.5).bfloat16()
    yc = ops.conv3x3(x, wc)
    assert yc.shape == (NI, 64, 64, C) and torch.equal(ops.conv3x3(x * 4, wc), yc * 4)
    xs = rn(NI, 32, 32, 640).bfloat16()
    w4 = rn(640, 640, 3, 3) * (9 * 640) ** -0.5
    wp, wph = w4.permute(0, 2, 3, 1).bfloat16().contiguous(), upsample_phase_weights(w4).bfloat16().contiguous()
    yu = ops.conv3x3(xs, wp, upsample=2, w_phases=wph)
    assert yu.shape == (NI, 64, 64, 640) and torch.equal(ops.conv3x3(xs * 2, wp, upsample=2, w_phases=wph), yu * 2)
    # attention: V constant per channel -> every output row equals it (softmax rows sum to 1)
    heads, D, NB = 8, 40, NBA
    qk = torch.zeros(NB, L, 2 * heads * 64 + C, device=DEV, dtype=torch.bfloat16)
    qk[:, :, :2 * heads * 64].view(NB, L, 2 * heads, 64)[..., :D] = rn(NB, L, 2 * heads, D).bfloat16()
    vconst = rn(C).bfloat16()
    vt = vconst.view(1, C, 1).expand(NB, C, L).contiguous()
    o = ops.self_attention_tc(qk, 0, heads * 64, vt, heads, D, D ** -0.5)
    assert float((o.float() - vconst.float()).abs().max()) < 2e-2 * float(vconst.float().abs().max())
    # norms: standardised outputs
    xg = (rn(2, RG, C) * 3 + 1).bfloat16()
    one, zero = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    gn = ops.groupnorm(xg, one, zero, 32, 1e-5, stat_batches=2).float().view(2, RG, 32, C // 32)
    assert float(gn.mean(dim=(1, 3)).abs().max()) < 1e-2 and float((gn.var(dim=(1, 3), unbiased=False) - 1).abs().max()) < 2e-2
    ln = ops.layernorm(xg.view(-1, C), one, zero).float()
    assert float(ln.mean(1).abs().max()) < 2e-2 and float((ln.var(1, unbiased=False) - 1).abs().max()) < 5e-2
    # temporal attention: each (clip, pixel, head) is independent -> permuting pixels permutes the output, bit for bit
    qkv = rn(2, 16, PX, 3 * C).bfloat16()
    perm = torch.randperm(PX, generator=g, device=DEV)
    assert torch.equal(ops.temporal_attention(qkv[:, :, perm].contiguous(), 8, D ** -0.5), ops.temporal_attention(qkv, 8, D ** -0.5)[:, :, perm])


def test_full_size_unet_and_pipeline_properties():
    """cfg2-size UNet forward: run-to-run bit reproducibility, independence of the CFG halves (changing the cond text changes only the
    cond prediction), hoisted ClipContext == per-call context; two DDIM steps + decode of all frames: finite frames in [0, 1],
    reproducible; per-frame VAE decode independent of its batch neighbours."""
    from followyourclick_b200 import AnimationPipeline, AutoencoderKL, DDIMScheduler, ops
    from followyourclick_b200.synth import synth_on_device_
    from tests.cfgs import SCHED_V
    unet = _full_unet()
    F_, hw = (4, 16) if SMALL else (16, 64)
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.zeros(2, F_, hw, hw, unet.input_channel_pad(), device=DEV, dtype=torch.bfloat16)
    x[..., :9] = torch.randn(1, F_, hw, hw, 9, generator=g, device=DEV).bfloat16()          # a CFG pair: identical inputs
    ctx = torch.randn(2, 77, 768, generator=g, device=DEV)
    kw = dict(fps_tensor=torch.tensor([2, 2], device=DEV), flow_control=torch.tensor([4, 4], device=DEV), use_fps_condition=True)
    t = torch.tensor(501, device=DEV)
    a = unet.forward_nfhwc(x, t, ctx, **kw).float()
    b = unet.forward_nfhwc(x, t, ctx, context=unet.prepare_context(ctx), **kw).float()
    assert a.shape == (2, F_, hw, hw, 4) and bool(torch.isfinite(a).all()) and torch.equal(a, b)
    ctx2 = ctx.clone()
    ctx2[1] += 0.5
    c = unet.forward_nfhwc(x, t, ctx2, **kw).float()
    assert torch.equal(a[0], c[0]) and not torch.equal(a[1], c[1])
    vae = synth_on_device_(AutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                                         up_block_types=("UpDecoderBlock2D",) * 4,
                                         block_out_channels=(32, 64, 128, 128) if SMALL else (128, 256, 512, 512),
                                         layers_per_block=1 if SMALL else 2, latent_channels=4, norm_num_groups=32).to(DEV).to(torch.bfloat16), seed=1)
    z = torch.randn(4, hw, hw, 4, generator=g, device=DEV).bfloat16()
    full, solo = vae.decode_nhwc(z).float(), vae.decode_nhwc(z[1:2].contiguous()).float()
    assert full.shape == (4, 8 * hw, 8 * hw, 3) and float((full[1:2] - solo).norm() / solo.norm()) < 1e-2     # (statistics chunking depends on the batch: last-bit bf16 differences)
    tok = type("Tok", (), dict(model_max_length=77, __call__=lambda self, p, **k: type("T", (), dict(
        input_ids=torch.zeros(1, 77, dtype=torch.long), attention_mask=torch.ones(1, 77, dtype=torch.long)))()))()
    emb = torch.randn(2, 77, 768, generator=torch.Generator().manual_seed(2))

    class Text(torch.nn.Module):
        calls, config = 0, type("C", (), {})()

        def forward(self, ids, attention_mask=None):
            Text.calls += 1
            return ((emb[1:2] if Text.calls % 2 == 1 else emb[0:1]).to(ids.device),)
    pipe = AnimationPipeline(vae=vae, text_encoder=Text(), tokenizer=tok, unet=unet, scheduler=DDIMScheduler(**SCHED_V))
    pipe.set_progress_bar_config(disable=True)
    lat = torch.randn(1, 4, F_, hw, hw, generator=torch.Generator().manual_seed(3))
    first = torch.randn(1, 4, hw, hw, generator=torch.Generator().manual_seed(4))
    if DEV != "cuda":
        pipe.use_cuda_graph = False
    run = lambda: pipe("p", negative_prompt="n", video_length=F_, height=8 * hw, width=8 * hw, num_inference_steps=2, guidance_scale=8.0,
                       latents=lat.clone(), use_first_frame_mask_condition_concat=True, first_image_latents=first, use_fps_condition=True,
                       fps_tensor=torch.tensor([2]), flow_control=torch.tensor([4])).videos
    v1, v2 = run(), run()
    assert v1.shape == (1, 3, F_, 8 * hw, 8 * hw) and v1.dtype == torch.float32 and not v1.is_cuda
    assert bool(torch.isfinite(v1).all()) and float(v1.min()) >= 0.0 and float(v1.max()) <= 1.0 and torch.equal(v1, v2)
