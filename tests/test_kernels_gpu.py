"""Per-kernel parity (GPU): every C-ABI entry point against a plain PyTorch fp32 reference of the same op.

Tolerances: fp32 kernels - rel-L2 <= 2e-5 (fp32 re-association only); bf16 storage with fp32 accumulation - rel-L2 <=
1.5e-2 against the fp32 reference fed the same bf16-rounded inputs (output rounding 2^-9 + operand rounding);
integer/elementwise fp32 glue (DDIM step, layout) - bit exact.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

F32_TOL = 2e-5
BF16_TOL = 1.5e-2


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def rnd(shape, seed, dtype=torch.float32, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def tol(dtype):
    return F32_TOL if dtype == torch.float32 else BF16_TOL


MODES = [(torch.float32, "simt"), (torch.bfloat16, "simt"), (torch.bfloat16, "tc")]


@pytest.fixture(autouse=True)
def _impl(cuda):
    from followyourclick_b200 import ops
    yield
    ops.set_impl("auto")


@pytest.mark.parametrize("dtype,impl", MODES)
@pytest.mark.parametrize("M,N,K", [(256, 320, 320), (300, 160, 64), (4096, 960, 320), (154, 640, 768), (128, 2560, 320),
                                   (2, 1280, 320), (1000, 48, 72)])
def test_gemm_bias_residual(dtype, impl, M, N, K):
    from followyourclick_b200 import ops
    if impl == "tc" and (M < 64 or N % 16 or K % 8):
        pytest.skip("shape not eligible for the tcgen05 path")
    ops.set_impl(impl)
    A, W = rnd((M, K), 1, dtype), rnd((N, K), 2, dtype, K ** -0.5)
    bias, res = rnd((N,), 3), rnd((M, N), 4, dtype)
    out = ops.gemm(A, W, bias=bias, residual=res, alpha=0.5)
    ref = 0.5 * (A.float() @ W.float().t()) + bias + res.float()
    assert out.dtype == dtype and rel(out, ref) < tol(dtype), rel(out, ref)


@pytest.mark.parametrize("dtype,impl", MODES)
def test_gemm_rowbias_and_f32_out(dtype, impl):
    from followyourclick_b200 import ops
    ops.set_impl(impl)
    M, N, K, rpg = 512, 320, 128, 128
    A, W, rb = rnd((M, K), 1, dtype), rnd((N, K), 2, dtype, K ** -0.5), rnd((M // rpg, N), 3)
    out = ops.gemm(A, W, rowbias=rb, rows_per_group=rpg, out_f32=True)
    ref = A.float() @ W.float().t() + rb.repeat_interleave(rpg, dim=0)
    assert out.dtype == torch.float32 and rel(out, ref) < (F32_TOL if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("dtype,impl", MODES)
def test_gemm_geglu(dtype, impl):
    from followyourclick_b200 import ops
    from followyourclick_b200.modeling import geglu_interleave
    ops.set_impl(impl)
    M, C = 384, 160
    x, w, b = rnd((M, C), 1, dtype), rnd((8 * C, C), 2, torch.float32, C ** -0.5), rnd((8 * C,), 3)
    wi, bi = geglu_interleave(w, b)
    out = ops.gemm(x, wi.to(dtype).contiguous(), bias=bi.contiguous(), geglu=True)
    h = x.float() @ w.to(dtype).float().t() + b
    a, g = h.chunk(2, dim=-1)
    ref = a * F.gelu(g)
    assert out.shape == (M, 4 * C) and rel(out, ref) < tol(dtype), rel(out, ref)


@pytest.mark.parametrize("M,N,K,res", [(128 * 9 - 30, 320, 512, True), (128 * 37, 1280, 1280, True), (128 * 16, 720, 640, False),
                                       (128 * 301, 320, 1280, True), (128 * 5 + 1, 48, 256, False)])
def test_gemm_tcgen05_cta_pairs(M, N, K, res, monkeypatch):
    """CTA-pair (cta_group::2) mode of the tcgen05 GEMM: odd m-block counts (phantom block in the last pair), ragged last N
    tile, multi-round persistent scheduling, residual prefetch across tiles."""
    from followyourclick_b200 import ops
    ops.set_impl("tc")
    dtype = torch.bfloat16
    A, W = rnd((M, K), 11, dtype), rnd((N, K), 12, dtype, K ** -0.5)
    bias, R = rnd((N,), 13), (rnd((M, N), 14, dtype) if res else None)
    ref = A.float() @ W.float().t() + bias + (R.float() if res else 0)
    outs = {}
    for mode in ("2", "0"):                     # 2 = pairs wherever legal, 0 = never (the dispatcher's own rule is FYC_TC_PAIR unset)
        monkeypatch.setenv("FYC_TC_PAIR", mode)
        outs[mode] = ops.gemm(A, W, bias=bias, residual=R)
        assert rel(outs[mode], ref) < tol(dtype), (mode, rel(outs[mode], ref))
        # bit-reproducible across launches (no atomics, fixed accumulation order)
        assert torch.equal(outs[mode], ops.gemm(A, W, bias=bias, residual=R))
    # same k order and fp32 accumulation in both modes: identical bits whenever both picked the same BN
    assert rel(outs["2"], outs["0"].float()) < 2e-3


def test_gemm_geglu_tcgen05_cta_pairs(monkeypatch):
    from followyourclick_b200 import ops
    from followyourclick_b200.modeling import geglu_interleave
    monkeypatch.setenv("FYC_TC_PAIR", "2")
    ops.set_impl("tc")
    dtype = torch.bfloat16
    M, C = 128 * 21, 640
    x, w, b = rnd((M, C), 1, dtype), rnd((8 * C, C), 2, torch.float32, C ** -0.5), rnd((8 * C,), 3)
    wi, bi = geglu_interleave(w, b)
    out = ops.gemm(x, wi.to(dtype).contiguous(), bias=bi.contiguous(), geglu=True)
    h = x.float() @ w.to(dtype).float().t() + b
    a, g = h.chunk(2, dim=-1)
    assert out.shape == (M, 4 * C) and rel(out, a * F.gelu(g)) < tol(dtype)

@pytest.mark.parametrize("dtype,impl", MODES)
def test_gemm_batched_scores_and_shared_A(dtype, impl):
    from followyourclick_b200 import ops
    ops.set_impl(impl)
    NB, HW, C = 3, 256, 128
    q, k = rnd((NB, HW, C), 1, dtype), rnd((NB, HW, C), 2, dtype)
    s = ops.gemm(q, k, alpha=C ** -0.5, out_f32=True)
    ref = torch.einsum("bik,bjk->bij", q.float(), k.float()) * C ** -0.5
    assert rel(s, ref) < (F32_TOL if dtype == torch.float32 else 4e-3)
    w = rnd((C, C), 3, dtype, C ** -0.5)
    vt = ops.gemm(w.unsqueeze(0).expand(NB, C, C), k)            # [NB, C, HW] = W @ k[n]^T
    assert rel(vt, torch.einsum("ck,bjk->bcj", w.float(), k.float())) < tol(dtype)


@pytest.mark.parametrize("dtype,impl", MODES)
@pytest.mark.parametrize("NB,H,W,Cin,Cout,stride,up", [(4, 16, 16, 64, 32, 1, 1), (2, 32, 32, 160, 160, 1, 1), (8, 8, 8, 320, 640, 1, 1),
                                                        (2, 16, 16, 64, 64, 2, 1), (2, 8, 8, 64, 48, 1, 2), (2, 16, 16, 9, 32, 1, 1),
                                                        (2, 16, 16, 32, 4, 1, 1), (1, 64, 64, 128, 128, 1, 1), (4, 12, 12, 64, 64, 1, 1)])
def test_conv3x3(dtype, impl, NB, H, W, Cin, Cout, stride, up):
    from followyourclick_b200 import ops
    if impl == "tc" and (Cin % 8 or Cout % 16):
        pytest.skip("shape not eligible for the tcgen05 path")
    ops.set_impl("auto" if impl == "tc" else impl)       # 'auto' so that ineligible patch shapes fall back instead of failing
    x = rnd((NB, H, W, Cin), 1, dtype)
    w = rnd((Cout, Cin, 3, 3), 2, torch.float32, (9 * Cin) ** -0.5)
    bias, temb = rnd((Cout,), 3), rnd((NB // 2 if NB > 1 else 1, Cout), 4)
    ipg = 2 if NB > 1 else 1
    wp = w.permute(0, 2, 3, 1).to(dtype).contiguous()
    Ho, Wo = (H * up + 2 - 3) // stride + 1, (W * up + 2 - 3) // stride + 1
    res = rnd((NB, Ho, Wo, Cout), 5, dtype)
    out = ops.conv3x3(x, wp, bias=bias, residual=res, rowbias=temb, images_per_group=ipg, stride=stride, upsample=up)
    xr = x.float().permute(0, 3, 1, 2)
    if up == 2:
        xr = F.interpolate(xr, scale_factor=2.0, mode="nearest")
    ref = F.conv2d(xr, w.to(dtype).float(), bias, stride=stride, padding=1)
    ref = ref + temb.repeat_interleave(ipg, dim=0)[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1) + res.float()
    assert out.shape == ref.shape and rel(out, ref) < tol(dtype), rel(out, ref)


@pytest.mark.parametrize("dtype,impl", MODES)
def test_conv3x3_rowbias_column_block_of_wider_table(dtype, impl):
    """fyc.h ld_rowbias: the row bias of a conv is a column block of the fused [B, sum Cout] time-embedding projection."""
    from followyourclick_b200 import ops
    ops.set_impl("auto" if impl == "tc" else impl)
    NB, H, W, Cin, Cout, ipg = 4, 16, 16, 64, 32, 2
    x = rnd((NB, H, W, Cin), 1, dtype)
    w = rnd((Cout, Cin, 3, 3), 2, torch.float32, (9 * Cin) ** -0.5)
    wp = w.permute(0, 2, 3, 1).to(dtype).contiguous()
    table = rnd((NB // ipg, 5 * Cout), 3)
    rb = table[:, 2 * Cout:3 * Cout]
    assert not rb.is_contiguous()
    out = ops.conv3x3(x, wp, rowbias=rb, images_per_group=ipg)
    same = ops.conv3x3(x, wp, rowbias=rb.contiguous(), images_per_group=ipg)
    assert torch.equal(out, same)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(dtype).float(), padding=1) + rb.repeat_interleave(ipg, 0)[:, :, None, None]
    assert rel(out, ref.permute(0, 2, 3, 1)) < tol(dtype)


@pytest.mark.parametrize("dtype,impl", MODES)
@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (3, 32, 32, 128, 160), (1, 8, 12, 24, 8)])
def test_conv3x3_stride2_bottom_right_pad(dtype, impl, NB, H, W, Cin, Cout):
    """pad_mode 1 = diffusers Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) + valid stride-2 conv (VAE encoder)."""
    from followyourclick_b200 import ops
    if impl == "tc" and (Cin % 8 or Cout % 16):
        pytest.skip("shape not eligible for the tcgen05 path")
    ops.set_impl("auto" if impl == "tc" else impl)
    x = rnd((NB, H, W, Cin), 1, dtype)
    w = rnd((Cout, Cin, 3, 3), 2, torch.float32, (9 * Cin) ** -0.5)
    bias = rnd((Cout,), 3)
    out = ops.conv3x3(x, w.permute(0, 2, 3, 1).to(dtype).contiguous(), bias=bias, stride=2, pad_mode=1)
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), w.to(dtype).float(), bias, stride=2).permute(0, 2, 3, 1)
    assert out.shape == ref.shape == (NB, H // 2, W // 2, Cout) and rel(out, ref) < tol(dtype), rel(out, ref)

@pytest.mark.parametrize("NB,H,W,Cin,Cout,stride", [(6, 16, 16, 128, 160, 1), (3, 32, 32, 64, 320, 1), (4, 32, 32, 128, 64, 2)])
def test_conv3x3_tcgen05_cta_pairs(NB, H, W, Cin, Cout, stride, monkeypatch):
    """Implicit-GEMM convolution on CTA pairs: patch tiles split over the two CTAs (incl. an odd number of m blocks and the
    stride-2 parity planes), row bias groups straddling a tile, residual."""
    from followyourclick_b200 import ops
    dtype = torch.bfloat16
    ops.set_impl("tc")
    x = rnd((NB, H, W, Cin), 1, dtype)
    w = rnd((Cout, Cin, 3, 3), 2, torch.float32, (9 * Cin) ** -0.5)
    bias, temb = rnd((Cout,), 3), rnd((NB, Cout), 4)
    wp = w.permute(0, 2, 3, 1).to(dtype).contiguous()
    Ho, Wo = H // stride, W // stride
    res = rnd((NB, Ho, Wo, Cout), 5, dtype)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(dtype).float(), bias, stride=stride, padding=1) + temb[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1) + res.float()
    for mode in ("2", "0"):
        monkeypatch.setenv("FYC_TC_PAIR", mode)
        out = ops.conv3x3(x, wp, bias=bias, residual=res, rowbias=temb, images_per_group=1, stride=stride)
        assert out.shape == ref.shape and rel(out, ref) < tol(dtype), (mode, rel(out, ref))

@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(2, 8, 8, 64, 48), (4, 16, 16, 128, 160), (32, 8, 8, 320, 320), (2, 32, 32, 64, 64),
                                             (3, 16, 8, 72, 32), (2, 64, 64, 128, 128)])
@pytest.mark.parametrize("pair", ["1", "2", "0"])
def test_conv3x3_upsample_phases_tcgen05(NB, H, W, Cin, Cout, pair, monkeypatch):
    """nearest-x2 + conv3x3 as four 2x2-tap implicit GEMMs written interleaved (fyc.h w_phases): against the fp32 PyTorch
    upsample + conv, and against the engine's own materialised-upsample path (same kernel, 9 taps)."""
    from followyourclick_b200 import _lib, ops
    from followyourclick_b200.modeling import upsample_phase_weights
    import ctypes as C
    dtype = torch.bfloat16
    ops.set_impl("auto")
    monkeypatch.setenv("FYC_TC_PAIR", pair)
    x = rnd((NB, H, W, Cin), 1, dtype)
    w = rnd((Cout, Cin, 3, 3), 2, torch.float32, (9 * Cin) ** -0.5)
    bias = rnd((Cout,), 3)
    wp = w.permute(0, 2, 3, 1).to(dtype).contiguous()
    wph = upsample_phase_weights(w).to(dtype).contiguous()
    a = _lib.ConvArgs(x.data_ptr(), wp.data_ptr(), x.data_ptr(), None, None, None, NB, H, W, Cin, Cout, 1, 2, 0, _lib.BF16, 0,
                      _lib.IMPL_AUTO, None, 0, 0, wph.data_ptr())
    assert _lib.lib().fyc_conv3x3_up2_eligible(C.byref(a)) == 1
    out = ops.conv3x3(x, wp, bias=bias, upsample=2, w_phases=wph)
    ref = F.conv2d(F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest"), w.to(dtype).float(), bias,
                   padding=1).permute(0, 2, 3, 1)
    assert out.shape == ref.shape == (NB, 2 * H, 2 * W, Cout) and rel(out, ref) < tol(dtype), rel(out, ref)
    old = ops.conv3x3(x, wp, bias=bias, upsample=2)                   # materialised upsample + 9-tap path
    assert rel(out, old) < 6e-3, rel(out, old)                        # two bf16 roundings of the same fp32 sums
    monkeypatch.setattr(ops, "use_up2_phases", False)
    assert torch.equal(ops.conv3x3(x, wp, bias=bias, upsample=2, w_phases=wph), old)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("NB,R,C,G,stat", [(2, 4 * 64, 160, 32, 2), (8, 64, 160, 32, 8), (2, 1024, 1920, 32, 2), (4, 16, 128, 32, 4),
                                            (3, 100, 36, 4, 3)])
@pytest.mark.parametrize("silu", [False, True])
@pytest.mark.parametrize("zigzag", ["1", "0"])
def test_groupnorm(cuda, dtype, NB, R, C, G, stat, silu, zigzag, monkeypatch):
    from followyourclick_b200 import ops
    monkeypatch.setenv("FYC_ZIGZAG", zigzag)       # back-to-front / front-to-back traversal of the statistics pass: same numbers
    x = (rnd((NB, R, C), 1) * 2 + 3).to(dtype)
    gamma, beta = rnd((C,), 2) + 1, rnd((C,), 3)
    out = ops.groupnorm(x, gamma, beta, G, 1e-5, silu=silu, stat_batches=stat)
    ref = F.group_norm(x.float().permute(0, 2, 1), G, gamma, beta, 1e-5)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 1)
    assert rel(out, ref) < (1e-5 if dtype == torch.float32 else 6e-3), rel(out, ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [160, 320, 640, 1280, 768])
def test_layernorm_with_pe(cuda, dtype, C, monkeypatch):
    from followyourclick_b200 import ops
    x0 = (rnd((2 * 4 * 16 + 3, C), 1) * 1.5 + 0.5).to(dtype)          # ragged row count: the last row block is partial
    g0, b0 = rnd((C,), 2) + 1, rnd((C,), 3)
    monkeypatch.setenv("FYC_ZIGZAG", "0")
    fwd = ops.layernorm(x0, g0, b0)
    monkeypatch.setenv("FYC_ZIGZAG", "1")
    assert torch.equal(ops.layernorm(x0, g0, b0), fwd)                  # traversal order does not change a single bit
    assert rel(fwd, F.layer_norm(x0.float(), (C,), g0, b0, 1e-5)) < (1e-5 if dtype == torch.float32 else 6e-3)
    Fr, HW = 4, 16
    x = (rnd((2 * Fr * HW, C), 1) * 1.5 + 0.5).to(dtype)
    gamma, beta, pe = rnd((C,), 2) + 1, rnd((C,), 3), rnd((24, C), 4)
    out = ops.layernorm(x, gamma, beta)
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    assert rel(out, ref) < (1e-5 if dtype == torch.float32 else 6e-3)
    out = ops.layernorm(x, gamma, beta, pe=pe, rows_per_frame=HW, frames=Fr)
    ref2 = (ref.view(2, Fr, HW, C) + pe[:Fr].view(1, Fr, 1, C)).view(-1, C)
    assert rel(out, ref2) < (1e-5 if dtype == torch.float32 else 6e-3)


def _mha_ref(q, k, v, heads, scale):
    B, Lq, C = q.shape
    d = C // heads
    qh = q.float().view(B, Lq, heads, d).transpose(1, 2)
    kh = k.float().view(k.shape[0], -1, heads, d).transpose(1, 2)
    vh = v.float().view(v.shape[0], -1, heads, d).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2) * scale).softmax(-1)
    return (s @ vh).transpose(1, 2).reshape(B, Lq, C)


@pytest.mark.parametrize("dtype,impl", MODES)
@pytest.mark.parametrize("heads,D,Lq,Lk", [(4, 40, 256, 256), (2, 80, 100, 77), (2, 160, 64, 64), (4, 40, 70, 81), (8, 40, 1024, 1024),
                                           (2, 160, 16, 4)])
def test_attention_self_and_cross(dtype, impl, heads, D, Lq, Lk):
    from followyourclick_b200 import ops
    ops.set_impl(impl)
    B, C = 4, heads * D
    qkv = rnd((B, Lq, 3 * C), 1, dtype)
    kv = rnd((B // 2, Lk, 2 * C), 2, dtype)
    scale = D ** -0.5
    if Lq == Lk:   # self attention on the fused qkv buffer (strided views)
        out = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads, scale)
        ref = _mha_ref(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads, scale)
        assert rel(out, ref) < tol(dtype), rel(out, ref)
    # cross attention: context shared by 2 consecutive batch entries, then an accumulated second pass (IP adapter)
    q = qkv[:, :, :C]
    out = ops.attention(q, kv[:, :, :C], kv[:, :, C:], heads, scale, kv_batch_div=2)
    kk, vv = kv[:, :, :C].repeat_interleave(2, 0), kv[:, :, C:].repeat_interleave(2, 0)
    ref = _mha_ref(q, kk, vv, heads, scale)
    assert rel(out, ref) < tol(dtype), rel(out, ref)
    T = 4
    ops.attention(q, kv[:, Lk - T:, :C], kv[:, Lk - T:, C:], heads, scale, out=out, out_alpha=0.7, accumulate=True, kv_batch_div=2)
    ref2 = ref + 0.7 * _mha_ref(q, kk[:, Lk - T:], vv[:, Lk - T:], heads, scale)
    assert rel(out, ref2) < tol(dtype) * 1.5, rel(out, ref2)


@pytest.mark.parametrize("heads,D,Lq,Lk,div", [(8, 40, 4096, 77, 2), (8, 80, 1000, 77, 4), (4, 64, 512, 128, 1), (8, 40, 1024, 16, 2),
                                               (2, 80, 256, 65, 1), (8, 40, 300, 4, 1)])
def test_attention_short_context_persistent_kernel(cuda, heads, D, Lq, Lk, div, monkeypatch):
    """Lk <= 128 (text / IP cross-attention): one CTA per (image, head) keeps K / V in shared memory and walks query tiles.  Same
    arithmetic as the generic kernel (agreement to one bf16 rounding), and within tolerance of the fp32 reference; ragged Lq / Lk, shared
    contexts (kv_batch_div) and the accumulated IP pass included."""
    from followyourclick_b200 import ops
    dtype = torch.bfloat16
    ops.set_impl("auto")
    B, C = 4, heads * D
    q = rnd((B, Lq, C), 1, dtype)
    kv = rnd((B // div, Lk, 2 * C), 2, dtype)
    scale = D ** -0.5
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("FYC_ATTN_SHORTK", mode)
        o = ops.attention(q, kv[:, :, :C], kv[:, :, C:], heads, scale, kv_batch_div=div)
        ops.attention(q, kv[:, :3, :C], kv[:, :3, C:], heads, scale, out=o, out_alpha=0.5, accumulate=True, kv_batch_div=div)
        outs[mode] = o
    # same arithmetic, separately compiled: the two kernels may differ by FMA contraction in the epilogue (observed for D = 40:
    # a last-bit bf16 difference on a few elements), never by more than one output rounding
    assert rel(outs["1"], outs["0"]) < 2e-3, rel(outs["1"], outs["0"])
    kk, vv = kv[:, :, :C].repeat_interleave(div, 0), kv[:, :, C:].repeat_interleave(div, 0)
    ref = _mha_ref(q, kk, vv, heads, scale) + 0.5 * _mha_ref(q, kk[:, :3], vv[:, :3], heads, scale)
    assert rel(outs["1"], ref) < tol(dtype) * 1.5, rel(outs["1"], ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("NB,R,C1,C2,stat,silu", [(2, 1024, 1280, 640, 2, True), (2, 256, 1280, 1280, 2, True), (4, 64, 320, 640, 4, False), (2, 4096, 640, 320, 2, True),
                                                  (2, 100, 64, 32, 1, True)])
def test_groupnorm_two_sources(cuda, dtype, NB, R, C1, C2, stat, silu):
    """fyc_groupnorm_concat: GroupNorm(+SiLU) of cat([x1, x2], channels) read in place (unet_blocks.py:763,885 -> resnet.py:240); groups that
    straddle the seam (1280 + 640 channels: 60-channel groups) included; equal to the kernel on the materialised concatenation."""
    import torch.nn.functional as Fn
    from followyourclick_b200 import ops
    x1, x2 = rnd((NB, R, C1), 1, dtype), (rnd((NB, R, C2), 2) * 1.5 + 0.3).to(dtype)
    gamma, beta = 1 + 0.1 * rnd((C1 + C2,), 3), 0.1 * rnd((C1 + C2,), 4)
    out = ops.groupnorm(x1, gamma, beta, 32, 1e-5, silu=silu, stat_batches=stat, x2=x2)
    cat = torch.cat([x1, x2], dim=-1)
    one = ops.groupnorm(cat, gamma, beta, 32, 1e-5, silu=silu, stat_batches=stat)
    assert out.shape == cat.shape and torch.equal(out, one)           # same arithmetic in the same order: bit-identical
    t = cat.float().reshape(stat, -1, C1 + C2).permute(0, 2, 1)
    ref = Fn.group_norm(t, 32, gamma, beta, 1e-5)
    ref = (Fn.silu(ref) if silu else ref).permute(0, 2, 1).reshape(cat.shape)
    assert rel(out, ref) < (2e-5 if dtype == torch.float32 else 6e-3), rel(out, ref)


@pytest.mark.parametrize("M,N,K1,K2,pair", [(8192, 1280, 1280, 1280, "1"), (32768, 640, 1280, 640, "1"), (4096, 320, 640, 320, "0"), (2048, 1280, 1280, 1280, "2"),
                                            (1000, 320, 320, 320, "0"), (512, 640, 640, 328, "0")])
def test_gemm_two_segment_k(cuda, M, N, K1, K2, pair, monkeypatch):
    """fyc_gemm_args.A2: the 1x1 shortcut over cat([x, skip]) with the K loop walking two tensor maps (resnet.py:286 after
    unet_blocks.py:763,885) - identical to the GEMM on the materialised concatenation (same k order, fp32 accumulation), CTA-pair mode
    and a ragged second segment (K2 % 64 != 0) included."""
    from followyourclick_b200 import ops
    monkeypatch.setenv("FYC_TC_PAIR", pair)
    ops.set_impl("tc")
    dt = torch.bfloat16
    a1, a2 = rnd((M, K1), 1, dt), rnd((M, K2), 2, dt)
    w, bias = rnd((N, K1 + K2), 3, dt, (K1 + K2) ** -0.5), rnd((N,), 4)
    out = ops.gemm(a1, w, bias=bias, A2=a2)
    cat = torch.cat([a1, a2], dim=1).contiguous()
    one = ops.gemm(cat, w, bias=bias)
    assert torch.equal(out, one)
    ref = cat.float() @ w.float().t() + bias
    assert rel(out, ref) < 4e-3, rel(out, ref)


def _ln_pack(w, gamma, beta, bias, dtype):
    """what UNet3DConditionModel._ln_fold packs: gamma-scaled, row-centred weight (rounded once), beta W^T + bias"""
    from followyourclick_b200 import ops
    return ops.ln_fold_weight(w, gamma, dtype), (w @ beta + (bias if bias is not None else 0)).contiguous()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,C", [(4096, 320), (1000, 640), (515, 1280), (300, 768), (64, 192)])
def test_layernorm_stats(cuda, dtype, M, C):
    from followyourclick_b200 import ops
    x = (rnd((M, C), 1) * 1.7 + rnd((M, 1), 2) * 3.0).to(dtype)       # per-row offsets: the mean term matters
    rs = ops.layernorm_stats(x)
    rstd = torch.rsqrt(x.float().var(dim=1, unbiased=False) + 1e-5)
    assert rs.shape == (M,) and rs.dtype == torch.float32 and rel(rs, rstd) < 1e-5


@pytest.mark.parametrize("M,N,K,rpg", [(4096, 960, 320, 0), (8192, 1344, 320, 0), (2048, 1920, 640, 256), (8192, 3840, 1280, 128), (1000, 320, 320, 0),
                                        (131072 // 8, 1344, 320, 0)])
def test_gemm_layernorm_fold(cuda, M, N, K, rpg):
    """FYC_EPI_LNFOLD: the GEMM on the RAW LayerNorm input + epilogue terms against LayerNorm -> GEMM in fp32 (attention.py:383,412;
    motion_module.py:261 incl. the position-table row bias), and against the unfused bf16 path it replaces (LN kernel -> bf16 -> GEMM)."""
    import torch.nn.functional as Fn
    from followyourclick_b200 import ops
    ops.set_impl("tc")
    dt = torch.bfloat16
    x = (rnd((M, K), 1) * 1.3 + rnd((M, 1), 2) * 8.0).to(dt)           # row means up to ~20 sigma: the balanced row sums keep the fold exact
    w = rnd((N, K), 3, torch.float32, K ** -0.5)
    gamma, beta, bias = 1 + 0.1 * rnd((K,), 4), 0.05 * rnd((K,), 5), 0.05 * rnd((N,), 6)
    wp, cb = _ln_pack(w, gamma, beta, bias, dt)
    rb = rnd((M // rpg, N), 7) if rpg else None
    out = ops.gemm(x, wp, bias=cb, rowbias=rb, rows_per_group=rpg, ln=ops.layernorm_stats(x))
    ref = Fn.layer_norm(x.float(), (K,), gamma, beta, 1e-5) @ w.t() + bias
    if rpg:
        ref = ref + rb.repeat_interleave(rpg, dim=0)
    unfused = ops.gemm(ops.layernorm(x, gamma.contiguous(), beta.contiguous()), w.to(dt).contiguous(), bias=bias.contiguous(), rowbias=rb, rows_per_group=rpg)
    e_f, e_u = rel(out, ref), rel(unfused, ref)
    assert e_f < 6e-3 and e_f < 1.5 * e_u + 1e-3, (e_f, e_u)          # no worse than LN -> bf16 -> GEMM (one rounding fewer on the activation)
    d = (out.float() - ref).abs()
    assert float(d.max()) < 2 ** -5 * float(ref.abs().max()), float(d.max())


@pytest.mark.parametrize("M,C", [(2048, 320), (4096, 640), (1024, 1280)])
def test_geglu_layernorm_fold(cuda, M, C):
    """norm3 / ff_norm folded into the GEGLU projection (value and gate columns both get rstd * acc + nrm * colsum + bias before a * gelu(g))"""
    import torch.nn.functional as Fn
    from followyourclick_b200 import ops
    from followyourclick_b200.modeling import geglu_interleave
    ops.set_impl("tc")
    dt = torch.bfloat16
    x = (rnd((M, C), 1) + rnd((M, 1), 2) * 2.0).to(dt)
    w, b = rnd((8 * C, C), 3, torch.float32, C ** -0.5), 0.05 * rnd((8 * C,), 4)
    gamma, beta = 1 + 0.1 * rnd((C,), 5), 0.05 * rnd((C,), 6)
    wp, cb = _ln_pack(w, gamma, beta, b, dt)
    wi, cbi = geglu_interleave(wp.float(), cb)
    out = ops.gemm(x, wi.to(dt).contiguous(), bias=cbi.contiguous(), geglu=True, ln=ops.layernorm_stats(x))
    h = Fn.layer_norm(x.float(), (C,), gamma, beta, 1e-5) @ w.t() + b
    a, g = h.chunk(2, dim=-1)
    ref = a * Fn.gelu(g)
    assert out.shape == (M, 4 * C) and rel(out, ref) < 8e-3, rel(out, ref)


@pytest.mark.parametrize("dtype,impl", MODES)
@pytest.mark.parametrize("heads,D,Lq,Lk,T,div", [(8, 40, 4096, 77, 16, 2), (8, 40, 1024, 77, 4, 4), (8, 80, 1024, 77, 16, 2), (4, 80, 100, 77, 4, 1),
                                                 (8, 160, 256, 77, 16, 2), (2, 160, 64, 77, 4, 1), (4, 40, 70, 64, 64, 1), (4, 40, 200, 150, 4, 2)])
def test_attention_fused_ip_second_context(dtype, impl, heads, D, Lq, Lk, T, div):
    """The IP-Adapter cross-attention as ONE launch (fyc.h second context): out = softmax(s q K_t^T) V_t + alpha2 softmax(s q K_i^T) V_i
    (animatediff/models/attention.py:92-120, ip_adapter/attention_processor.py:137-168) against the fp32 two-softmax reference and
    against the two-launch accumulate form it replaces.  Head dims of all UNet levels, T = 4 (vanilla) / 16 (plus), ragged Lq, shared
    contexts, a context that is too long for the resident kernel (150 keys: the dispatcher's two-pass route)."""
    from followyourclick_b200 import ops
    ops.set_impl(impl)
    B, C = 4, heads * D
    q = rnd((B, Lq, C), 1, dtype)
    kv, kvi = rnd((B // div, Lk + T, 2 * C), 2, dtype), rnd((B // div, Lk + T, 2 * C), 3, dtype)
    scale, a2 = D ** -0.5, 0.6
    k, v, k2, v2 = kv[:, :Lk, :C], kv[:, :Lk, C:], kvi[:, Lk:, :C], kvi[:, Lk:, C:]         # strided views, like the UNet's ClipContext
    out = ops.attention(q, k, v, heads, scale, kv_batch_div=div, k2=k2, v2=v2, alpha2=a2)
    rep = lambda t: t.repeat_interleave(div, 0)
    ref = _mha_ref(q, rep(k), rep(v), heads, scale) + a2 * _mha_ref(q, rep(k2), rep(v2), heads, scale)
    e = (out.float() - ref)
    assert rel(out, ref) < tol(dtype), rel(out, ref)
    assert float(e.abs().max()) < (1e-4 if dtype == torch.float32 else 2 ** -6 * float(ref.abs().max())), float(e.abs().max())
    two = ops.attention(q, k, v, heads, scale, kv_batch_div=div)
    ops.attention(q, k2, v2, heads, scale, out=two, out_alpha=a2, accumulate=True, kv_batch_div=div)
    assert rel(out, two) < (1e-6 if dtype == torch.float32 else 4e-3), rel(out, two)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,Fr,HW,heads,D", [(2, 4, 64, 4, 40), (2, 16, 16, 8, 40), (1, 8, 16, 4, 80), (2, 16, 4, 8, 160), (1, 24, 9, 2, 40),
                                             (1, 32, 4, 4, 160)])
def test_temporal_attention(cuda, dtype, B, Fr, HW, heads, D):
    from followyourclick_b200 import ops
    C = heads * D
    qkv = rnd((B, Fr, HW, 3 * C), 1, dtype)
    out = ops.temporal_attention(qkv, heads, D ** -0.5)
    t = qkv.float().permute(0, 2, 1, 3).reshape(B * HW, Fr, 3 * C)          # (b d) f c
    ref = _mha_ref(t[:, :, :C], t[:, :, C:2 * C], t[:, :, 2 * C:], heads, D ** -0.5)
    ref = ref.view(B, HW, Fr, C).permute(0, 2, 1, 3)
    assert rel(out, ref) < tol(dtype), rel(out, ref)


@pytest.mark.parametrize("L,heads,NB,gain", [(128, 2, 2, 1.0), (256, 4, 3, 1.0), (1024, 8, 2, 1.0), (512, 2, 1, 3.0)])
def test_self_attention_tcgen05(cuda, L, heads, NB, gain):
    """tcgen05 attention (S/O in TMEM) against the fp32 reference; gain 3 produces logits large enough to exercise the
    lazy O-rescale path (running max growing by more than 2^8 between key tiles)."""
    from followyourclick_b200 import ops
    if not ops.self_attention_tc_ok(torch.bfloat16, L, 40):
        pytest.skip("tcgen05 path unavailable")
    D, C = 40, heads * 40
    q = rnd((NB, L, heads, D), 1, torch.bfloat16, gain)
    k = rnd((NB, L, heads, D), 2, torch.bfloat16, gain)
    v = rnd((NB, L, C), 3, torch.bfloat16)
    if gain > 1:   # make the maximum grow along the key axis so later tiles raise the running max
        k = (k.float() * torch.linspace(0.2, 1.5, L, device="cuda").view(1, L, 1, 1)).bfloat16()
    qk = torch.zeros(NB, L, 2 * heads * 64 + C, dtype=torch.bfloat16, device="cuda")
    qk[:, :, :heads * 64].view(NB, L, heads, 64)[..., :D] = q
    qk[:, :, heads * 64:2 * heads * 64].view(NB, L, heads, 64)[..., :D] = k
    qk[:, :, 2 * heads * 64:] = v
    vt = ops.transpose_tokens(qk, 2 * heads * 64, C)
    assert torch.equal(vt, v.transpose(1, 2).contiguous())
    out = ops.self_attention_tc(qk, 0, heads * 64, vt, heads, D, D ** -0.5)
    ref = _mha_ref(q.reshape(NB, L, C), k.reshape(NB, L, C), v, heads, D ** -0.5)
    assert rel(out, ref) < BF16_TOL, rel(out, ref)
    # agreement with the mma.sync flash kernel on the same operands
    o2 = ops.attention(q.reshape(NB, L, C), k.reshape(NB, L, C), v, heads, D ** -0.5)
    assert rel(out, o2) < BF16_TOL


@pytest.mark.parametrize("L,heads,NB,gain", [(256, 2, 2, 1.0), (1024, 8, 3, 1.0), (512, 4, 1, 3.0), (2304, 8, 1, 1.0)])
def test_self_attention_tcgen05_d80(cuda, L, heads, NB, gain):
    """head dim 80 (level-1 attn1) on tcgen05: UNPADDED heads in the fused [q | k | v] buffer - the second 64-column atom of a head
    overlaps the next head (or, for the last k head, the v block), whose columns must not leak into the scores; against fp32
    softmax(q k^T / sqrt(80)) v on the same bf16 operands, and against the mma.sync kernel it replaces.  2304 = the cfg5 level-1 length."""
    from followyourclick_b200 import ops
    D = 80
    C = heads * D
    qkv = rnd((NB, L, 3 * C), 1, torch.bfloat16)
    qkv[:, :, :C] *= gain
    vt = ops.transpose_tokens(qkv, 2 * C, C)
    out = ops.self_attention_tc_d80(qkv, 0, C, vt, heads, D ** -0.5)
    ref = _mha_ref(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads, D ** -0.5)
    old = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], heads, D ** -0.5)
    e = (out.float() - ref)
    assert rel(out, ref) < BF16_TOL and float(e.abs().max()) < 2 ** -6 * float(ref.abs().max()) + 1e-3, (rel(out, ref), float(e.abs().max()))
    assert rel(out, old) < 1e-2, rel(out, old)


def _cross_pack(k, v, heads, D, Lpad):
    """test-side packing of a context for ops.cross_attention_tc: keys zero-padded to Lpad rows, heads to DKP columns; V transposed"""
    from followyourclick_b200 import ops
    NBc, Lk, C = k.shape
    dkp = ops.cross_dkp(D)
    kp = torch.zeros((NBc, Lpad, heads * dkp), dtype=k.dtype, device=k.device)
    kp.view(NBc, Lpad, heads, dkp)[:, :Lk, :, :D] = k.view(NBc, Lk, heads, D)
    vt = torch.zeros((NBc, C, Lpad), dtype=v.dtype, device=v.device)
    vt[:, :, :Lk] = v.transpose(1, 2)
    return kp, vt.contiguous()


@pytest.mark.parametrize("heads,D,Lq,Lk,T,div", [(8, 40, 4096, 77, 16, 2), (8, 40, 1024, 77, 4, 4), (8, 80, 1024, 77, 16, 2), (4, 80, 1000, 77, 0, 1),
                                                 (4, 40, 300, 80, 16, 1), (8, 40, 128, 5, 1, 2), (2, 80, 130, 77, 4, 1), (8, 80, 2304, 77, 4, 2)])
def test_cross_attention_tcgen05(cuda, heads, D, Lq, Lk, T, div):
    """tcgen05 cross-attention with a resident short context (fyc_cross_attention_tc): text keys + optional image keys in one launch,
    UNPADDED q heads (the 64-column box of a D = 40 head reads 24 foreign columns that meet zero key columns), padding keys masked,
    ragged query counts, shared contexts - against the fp32 two-softmax reference and the mma.sync kernel it replaces."""
    from followyourclick_b200 import ops
    dt = torch.bfloat16
    B, C = 4, heads * D
    q = rnd((B, Lq, C), 1, dt)
    kt, vtx = rnd((B // div, Lk, C), 2, dt), rnd((B // div, Lk, C), 3, dt)
    scale, a1, a2 = D ** -0.5, 1.0, 0.6
    kp, vt = _cross_pack(kt, vtx, heads, D, ops.CROSS_LK)
    k2 = vt2 = ki = vi = None
    if T:
        ki, vi = rnd((B // div, T, C), 4, dt), rnd((B // div, T, C), 5, dt)
        k2, vt2 = _cross_pack(ki, vi, heads, D, ops.CROSS_LK2)
    out = torch.full((B, Lq, C), float("nan"), dtype=dt, device="cuda")
    ops.cross_attention_tc(q, kp, vt, heads, D, scale, Lk, out, k2=k2, vt2=vt2, Lk2=T, out_alpha=a1, alpha2=a2, kv_batch_div=div)
    rep = lambda t: t.repeat_interleave(div, 0)
    ref = a1 * _mha_ref(q, rep(kt), rep(vtx), heads, scale)
    if T:
        ref = ref + a2 * _mha_ref(q, rep(ki), rep(vi), heads, scale)
    assert bool(torch.isfinite(out).all())
    e = (out.float() - ref)
    assert rel(out, ref) < BF16_TOL and float(e.abs().max()) < 2 ** -6 * float(ref.abs().max()) + 1e-3, (rel(out, ref), float(e.abs().max()))
    old = ops.attention(q, kt, vtx, heads, scale, kv_batch_div=div, k2=ki, v2=vi, alpha2=a2)
    assert rel(out, old) < 1e-2, rel(out, old)


def test_softmax_rows_and_misc(cuda):
    from followyourclick_b200 import ops
    s = rnd((64, 300), 1) * 4
    assert rel(ops.softmax_rows(s, torch.float32), s.softmax(-1)) < 1e-6
    assert rel(ops.softmax_rows(s, torch.bfloat16), s.softmax(-1)) < 5e-3
    x = rnd((3, 5, 7, 16), 2)
    assert torch.equal(ops.upsample_nearest2x(x), x.repeat_interleave(2, 1).repeat_interleave(2, 2))
    xb = x.bfloat16()
    assert torch.equal(ops.upsample_nearest2x(xb), xb.repeat_interleave(2, 1).repeat_interleave(2, 2))
    a, b = rnd((10, 24), 3), rnd((10, 40), 4)
    assert torch.equal(ops.concat_channels(a, b), torch.cat([a, b], -1))
    assert torch.equal(ops.concat_channels(a.bfloat16(), b.bfloat16()), torch.cat([a, b], -1).bfloat16())
    assert rel(ops.silu(a), F.silu(a)) < 1e-6
    v = rnd((2, 4, 3, 5, 6), 5)
    nf = ops.ncfhw_to_nfhwc(v, torch.float32)
    assert torch.equal(nf, v.permute(0, 2, 3, 4, 1).contiguous())
    assert torch.equal(ops.nfhwc_to_ncfhw(nf), v)
    assert torch.equal(ops.ncfhw_to_nfhwc(v, torch.float32, scale=1 / 0.18215), (1 / 0.18215 * v).permute(0, 2, 3, 4, 1).contiguous())


def test_timestep_embed_matches_reference_formula(cuda):
    from followyourclick_b200 import ops
    dim, half = 320, 160
    t = torch.tensor([961, 501, 1, 2, 4], dtype=torch.int64)
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32) / half
    freqs = torch.exp(exponent)
    emb = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(emb), torch.sin(emb)], -1)        # flip_sin_to_cos=True
    out = ops.timestep_embed(t.cuda(), freqs.cuda(), True)
    assert float((out.cpu() - ref).abs().max()) < 2e-6


def test_build_unet_input_and_finalize(cuda):
    from followyourclick_b200 import ops
    from oracle.ref_pipeline import build_unet_input
    b, f, h, w = 2, 3, 4, 5
    lat, first = rnd((b, 4, f, h, w), 1), rnd((b, 4, h, w), 2)
    mask = (rnd((b, 1, 1, h, w), 3) * 2)
    ref = build_unet_input(lat.cpu(), first.cpu(), mask.cpu(), True).permute(0, 2, 3, 4, 1)
    out = ops.build_unet_input(lat, mask[:, :, 0].contiguous(), first, 2, torch.float32)
    assert torch.equal(out.cpu(), ref.contiguous())
    ref0 = build_unet_input(lat.cpu(), first.cpu(), None, True).permute(0, 2, 3, 4, 1)
    assert torch.equal(ops.build_unet_input(lat, None, first, 2, torch.float32).cpu(), ref0.contiguous())
    assert torch.equal(ops.build_unet_input(lat, None, None, 1, torch.float32), lat.permute(0, 2, 3, 4, 1).contiguous())
    x = rnd((b * f, 8, 8, 3), 4) * 2
    vid = ops.frames_finalize(x, b, f)
    refv = (x.view(b, f, 8, 8, 3).permute(0, 4, 1, 2, 3) / 2 + 0.5).clamp(0, 1)
    assert torch.equal(vid, refv.contiguous())
    # channel-strided readers: the first 3 / 4 of 16 channels (zero-padded tcgen05 heads)
    wide = rnd((b * f, 8, 8, 16), 5) * 2
    assert torch.equal(ops.frames_finalize(wide[..., :3], b, f), (wide[..., :3].reshape(b, f, 8, 8, 3).permute(0, 4, 1, 2, 3) / 2 + 0.5).clamp(0, 1))
    wb = wide.to(torch.bfloat16).view(b, f, 8, 8, 16)
    assert torch.equal(ops.nfhwc_to_ncfhw(wb[..., :4]), wb[..., :4].float().permute(0, 4, 1, 2, 3).contiguous())
    assert torch.equal(ops.nfhwc_to_ncfhw(wb), wb.float().permute(0, 4, 1, 2, 3).contiguous())


@pytest.mark.parametrize("NB,H,W,Cin,cout,f32", [(32, 16, 16, 320, 4, True), (2, 64, 64, 128, 3, False), (4, 32, 32, 64, 4, True)])
def test_conv_head_zero_padded_to_16_on_tcgen05(cuda, NB, H, W, Cin, cout, f32):
    """conv_out heads (UNet 320 -> 4 fp32, VAE 128 -> 3): N = 16 tiles on the tcgen05 path, first `cout` channels == the conv."""
    from followyourclick_b200 import ops
    dtype = torch.bfloat16
    ops.set_impl("tc")
    x = rnd((NB, H, W, Cin), 1, dtype)
    w = rnd((cout, Cin, 3, 3), 2, torch.float32, (9 * Cin) ** -0.5)
    bias = rnd((cout,), 3)
    wp = torch.zeros(16, 3, 3, Cin, dtype=dtype, device="cuda")
    wp[:cout] = w.permute(0, 2, 3, 1).to(dtype)
    bp = torch.zeros(16, device="cuda")
    bp[:cout] = bias
    out = ops.conv3x3(x, wp, bias=bp, out_f32=f32)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(dtype).float(), bias, padding=1).permute(0, 2, 3, 1)
    assert out.dtype == (torch.float32 if f32 else dtype) and out.shape == (NB, H, W, 16)
    assert rel(out[..., :cout], ref) < (2e-3 if f32 else tol(dtype)), rel(out[..., :cout], ref)
    assert float(out[..., cout:].float().abs().max()) == 0.0
    A, Wm = rnd((300, 72), 4, dtype), rnd((16, 72), 5, dtype, 72 ** -0.5)
    assert rel(ops.gemm(A, Wm), A.float() @ Wm.float().t()) < tol(dtype)


def test_cfg_ddim_step_bit_exact_vs_golden(cuda):
    """DDIMScheduler.step against the reference-generated fixture, bit for bit (both prediction types, eta > 0)."""
    import numpy as np
    from followyourclick_b200 import DDIMScheduler
    from tests.cfgs import SCHED_EPS, SCHED_V
    from tests.engine_helpers import golden
    g = golden("ddim.npz")
    for name, cfg in (("v", SCHED_V), ("eps", SCHED_EPS)):
        sch = DDIMScheduler(**cfg)
        x, v = torch.from_numpy(g[f"{name}_x"]).cuda(), torch.from_numpy(g[f"{name}_v"]).cuda()
        for n in (4, 25, 50):
            sch.set_timesteps(n, device="cuda")
            assert np.array_equal(sch.timesteps.cpu().numpy(), g[f"{name}_timesteps_{n}"])
            ts = sch._timesteps_host
            for t in (ts[0], ts[len(ts) // 2], ts[-1]):
                out = sch.step(v, t, x).prev_sample
                assert np.array_equal(out.cpu().numpy(), g[f"{name}_step_{n}_{t}"]), (name, n, t)
        sch.set_timesteps(25, device="cuda")
        out = sch.step(v, 481, x, eta=0.5, variance_noise=torch.from_numpy(g[f"{name}_noise"]).cuda()).prev_sample
        assert np.array_equal(out.cpu().numpy(), g[f"{name}_step_eta0.5_25_481"])
    # fused CFG: u + s (c - u) then step == two-stage reference arithmetic
    sch = DDIMScheduler(**SCHED_V)
    sch.set_timesteps(25, device="cuda")
    u, c, x = rnd((1, 4, 4, 8, 8), 1), rnd((1, 4, 4, 8, 8), 2), rnd((1, 4, 4, 8, 8), 3)
    fused = sch.step_cfg(torch.cat([u, c]), 481, x, 8.0)
    two = sch.step(u + 8.0 * (c - u), 481, x).prev_sample
    assert torch.equal(fused, two)


@pytest.mark.parametrize("b,nrow,rescale", [(1, 6, False), (3, 6, False), (7, 4, False), (4, 2, True)])
def test_video_grid_u8_bit_exact(cuda, b, nrow, rescale):
    """fyc_video_grid_u8 == save_videos_grid's make_grid + (x * 255).astype(uint8) (oracle/ref_util.py), byte for byte."""
    import numpy as np
    from followyourclick_b200 import ops
    from oracle import ref_util
    g = torch.Generator().manual_seed(9)
    v = torch.rand(b, 3, 3, 10, 12, generator=g)
    if rescale:
        v = v * 2 - 1
    v[0, 0, 0, 0, 0], v[0, 1, 0, 0, 0] = 1.0, (-1.0 if rescale else 0.0)
    out = ops.video_grid_u8(v.cuda(), nrow=nrow, rescale=rescale).cpu().numpy()
    ref = np.stack(ref_util.video_frames_uint8(v, rescale=rescale, n_rows=nrow))
    assert out.shape == ref.shape == ops.video_grid_shape(b, 3, 10, 12, nrow) and np.array_equal(out, ref)


def test_save_videos_grid_writes_a_gif(cuda, tmp_path):
    from PIL import Image
    from followyourclick_b200.util import save_videos_grid
    v = torch.rand(2, 3, 4, 16, 16, generator=torch.Generator().manual_seed(1))
    p = save_videos_grid(v.cuda(), str(tmp_path / "a" / "s.gif"), n_rows=2)
    im = Image.open(p)
    assert im.n_frames == 4 and im.size == (2 * 18 + 2, 18 + 2)
    save_videos_grid(v, str(tmp_path / "b.gif"))           # a host tensor (the reference API's return type) also works
