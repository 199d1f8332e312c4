// Spatial self-attention on 5th-gen tensor cores (tcgen05) for the long-sequence, small-head case that dominates the
// UNet (level 0: 4096 tokens, 8 heads x 40 dims, 256 (image, head) problems per forward, 5 forwards-worth per step).
//
// One CTA = 128 query rows of one (image, head).  Per 128-key tile:
//   MMA warp      S = Q K^T       tcgen05.mma 128 x 128 x 64 (head dim zero-padded to 64 by the projection GEMM), fp32 S in TMEM
//   softmax warps tcgen05.ld one S row per thread (128 fp32 registers), running max with lazy rescale (O is touched only when
//                 the max grows by > 2^8), p = ex2(s * scale*log2e - m), row sum; P (bf16) written to shared memory in the
//                 128B-swizzled K-major layout the tensor core reads; O rescaled in TMEM with tcgen05.ld/st when needed
//   MMA warp      O += P V        tcgen05.mma 128 x 48 x 128 with V^T (keys contiguous) as the K-major B operand, fp32 O in TMEM
// S is double-buffered in TMEM and P in shared memory, so QK^T of tile j+1 overlaps the softmax of tile j; K and V^T tiles
// arrive through two 2-stage TMA rings.  The score matrix never exists outside TMEM/registers.
// Operand prerequisites (prepared by the host side once per layer call, see unet.py::_transformer):
//   q, k : [NB, L, heads * 64] bf16 (zero columns 40..63 per head come for free from zero rows in the packed projection weight)
//   v^T  : [NB, heads * D, L] bf16 (fyc_transpose_tokens)
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int BQ = 128, BKV = 128, DPAD = 64, DV = 48;        // DV: PV accumulator columns (D = 40 padded to a UMMA N multiple of 16)
constexpr int Q_BYTES = BQ * DPAD * 2;                         // 16 KB
constexpr int K_BYTES = BKV * DPAD * 2;                        // 16 KB per stage
constexpr int VBOX_BYTES = DV * 64 * 2;                        // 6 KB: one box = 48 rows (d) x 64 keys
constexpr int V_BYTES = 2 * VBOX_BYTES;                        // 12 KB per stage (keys 0-63 | 64-127)
constexpr int PHALF_BYTES = BQ * 64 * 2;                       // 16 KB: 128 rows x 64 keys
constexpr int P_BYTES = 2 * PHALF_BYTES;                       // 32 KB per buffer
constexpr int OFF_K = Q_BYTES, OFF_V = OFF_K + 2 * K_BYTES, OFF_P = OFF_V + 2 * V_BYTES, OFF_BAR = OFF_P + 2 * P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int NTHREADS = 192;
constexpr int TM_S0 = 0, TM_S1 = 128, TM_O = 256;              // TMEM column map

struct AttnTcParams {
  bf16* out; int64_t ldo, bso;      // out[n, token, h*D + d]
  int L, heads, D;
  float scale_log2e, out_alpha;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {     // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32 pairs (FFMA2 / FADD2): one issue slot for two elements
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 p, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// 2^x for a pair on the FMA pipe (Cody-Waite: n = round(x), cubic minimax of 2^f on [-0.5, 0.5], max rel. error 7.5e-5 - fifty
// times below the bf16 rounding P gets anyway; exponent patched in with an integer add).  x in [-126, 126].
// The MUFU unit does 16 ex2 per clock per SM and is what bounds the softmax (128 x 128 exps per 4 MMAs); moving 3 of every 8
// 8-element chunks here balances the two pipes.
__device__ __forceinline__ f32x2 exp2_poly2(f32x2 x) {
  const f32x2 magic = pk2(12582912.0f, 12582912.0f), nmagic = pk2(-12582912.0f, -12582912.0f), m1 = pk2(-1.0f, -1.0f);
  const f32x2 c0 = pk2(0.9999280571937561f, 0.9999280571937561f), c1 = pk2(0.6932609677314758f, 0.6932609677314758f);
  const f32x2 c2 = pk2(0.2426111251115799f, 0.2426111251115799f), c3 = pk2(0.05517164617776871f, 0.05517164617776871f);
  const f32x2 t = add2(x, magic);                 // low mantissa bits of t = round-to-nearest integer of x
  const f32x2 f = fma2(add2(t, nmagic), m1, x);   // x - n  in [-0.5, 0.5]
  f32x2 p = fma2(f, c3, c2);
  p = fma2(p, f, c1);
  p = fma2(p, f, c0);
  float t0, t1, p0, p1; upk2(t, t0, t1); upk2(p, p0, p1);
  return pk2(__uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23)), __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23)));
}

__global__ void __launch_bounds__(NTHREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_vt, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024 - (raw & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;            // [1]
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;        // [2]
  uint64_t* v_empty = bars + 7;       // [2]
  uint64_t* s_full = bars + 9;        // [2]
  uint64_t* s_empty = bars + 11;      // [2]
  uint64_t* p_full = bars + 13;       // [2]
  uint64_t* p_empty = bars + 15;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int T = p.L / BKV;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_vt)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 128); mbar_init(&p_full[i], 128); mbar_init(&p_empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);
      tma_load_4d(&map_q, q_full, smem, 0, qt * BQ, h, n);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], K_BYTES);
        tma_load_4d(&map_k, &k_full[st], smem + OFF_K + st * K_BYTES, 0, j * BKV, h, n);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], V_BYTES);
        uint8_t* sv = smem + OFF_V + st * V_BYTES;
        tma_load_3d(&map_vt, &v_full[st], sv, j * BKV, h * p.D, n);
        tma_load_3d(&map_vt, &v_full[st], sv + VBOX_BYTES, j * BKV + 64, h * p.D, n);
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BKV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const int kq = (p.D + 15) >> 4;      // k-steps of QK^T that hold data: q / k columns D..63 are zero, a 16-column step of zeros is skipped (D = 40: 3 of 4)
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint64_t q_desc = sw128_desc(smem_u32(smem));
      auto issue_s = [&](int j) {
        const int b = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_full[b], ph);
        mbar_wait(&s_empty[b], ph ^ 1);
        tc_fence_after();
        const uint64_t k_desc = sw128_desc(smem_u32(smem + OFF_K + b * K_BYTES));
#pragma unroll
        for (int kk = 0; kk < kq; ++kk)
          umma(tmem_base + (b ? TM_S1 : TM_S0), q_desc + 2 * kk, k_desc + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
        tc_commit(&k_empty[b]);
        tc_commit(&s_full[b]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        const int b = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&v_full[b], ph);
        mbar_wait(&p_full[b], ph);
        tc_fence_after();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint64_t a_desc = sw128_desc(smem_u32(smem + OFF_P + b * P_BYTES + hh * PHALF_BYTES));
          const uint64_t b_desc = sw128_desc(smem_u32(smem + OFF_V + b * V_BYTES + hh * VBOX_BYTES));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma(tmem_base + TM_O, a_desc + 2 * kk, b_desc + 2 * kk, idesc_o, (j > 0 || hh > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(&v_empty[b]);
        tc_commit(&p_empty[b]);
      }
    }
  } else {
    // ================================================================== softmax / correction / epilogue (warps 2..5)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                       // query row of the tile owned by this thread
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale_log2e;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < T; ++j) {
      const int b = j & 1; const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[b], ph);
      tc_fence_after();
      uint32_t sr[128];
      const uint32_t s_addr = tmem_base + lane_base + (b ? TM_S1 : TM_S0);
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(s_addr + c * 32, sr + c * 32);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&s_empty[b]);                               // S_b may be overwritten by tile j+2
      float mt = __uint_as_float(sr[0]);
#pragma unroll
      for (int i = 1; i < 128; ++i) mt = fmaxf(mt, __uint_as_float(sr[i]));
      // lazy rescale: keep the stale maximum unless it would let p exceed 2^8 (bf16/fp32 have the range to absorb it)
      float factor = 1.0f;
      bool need = false;
      if (j == 0) {
        m_used = mt;
      } else if ((mt - m_used) * sl2 > 8.0f) {
        factor = ex2((m_used - mt) * sl2);
        m_used = mt;
        need = true;
      }
      l *= factor;
      const float nb = -m_used * sl2;
      mbar_wait(&p_empty[b], ph ^ 1);                         // P_b free (PV of tile j-2 retired)
      uint8_t* prow = smem + OFF_P + b * P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {                          // 16 chunks of 8 keys = 16 bytes of bf16
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { e[i] = ex2(fmaf(__uint_as_float(sr[c * 8 + i]), sl2, nb)); sum += e[i]; }
        uint4 v;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(e[0], e[1]), t1 = __floats2bfloat162_rn(e[2], e[3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(e[4], e[5]), t3 = __floats2bfloat162_rn(e[6], e[7]);
        v.x = *reinterpret_cast<uint32_t*>(&t0); v.y = *reinterpret_cast<uint32_t*>(&t1);
        v.z = *reinterpret_cast<uint32_t*>(&t2); v.w = *reinterpret_cast<uint32_t*>(&t3);
        const int half = c >> 3, cc = c & 7;                  // 64-key half tile, 16-byte chunk inside the 128-byte row
        *reinterpret_cast<uint4*>(prow + half * PHALF_BYTES + ((cc ^ (r & 7)) << 4)) = v;
      }
      l += sum;
      if (__any_sync(0xffffffffu, need)) {
        // O must be quiescent: PV of tile j-1 retired (its commit completes p_empty[b^1] for that tile's phase)
        mbar_wait(&p_empty[b ^ 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t o_addr = tmem_base + lane_base + TM_O;
#pragma unroll
        for (int c = 0; c < DV / 16; ++c) {
          uint32_t orr[16];
          tmem_ld16(o_addr + c * 16, orr);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * factor);
          tmem_st16(o_addr + c * 16, orr);
        }
        tmem_wait_st();
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // P (generic-proxy stores) -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(&p_full[b]);
    }
    // ---- epilogue: O / l
    mbar_wait(&p_empty[(T - 1) & 1], ((T - 1) >> 1) & 1);
    tc_fence_after();
    const float inv = p.out_alpha / l;
    bf16* orow = p.out + (int64_t)n * p.bso + ((int64_t)qt * BQ + r) * p.ldo + h * p.D;
    const uint32_t o_addr = tmem_base + lane_base + TM_O;
#pragma unroll
    for (int c = 0; c < DV / 16; ++c) {
      uint32_t orr[16];
      tmem_ld16(o_addr + c * 16, orr);
      tmem_wait_ld();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(orr[i]) * inv;
      if (c * 16 + 8 <= p.D) Vec8<bf16>::store(orow + c * 16, f);
      if (c * 16 + 16 <= p.D) Vec8<bf16>::store(orow + c * 16 + 8, f + 8);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two query tiles per CTA ("ping-pong"): 256 query rows, two softmax warpgroups (warps 2-5 -> rows 0-127, warps 6-9 -> rows
// 128-255), each with its own S / O accumulators in TMEM and its own P buffer.  With one softmax warp per scheduler the
// single-tile kernel above is bound by that warp's serial instruction stream (2230 clk per 128x128 tile, measured); two warps
// per scheduler overlap each other's MUFU / FMA / shared-memory latencies, and every K / V^T tile is loaded once for 256 rows.
constexpr int G2_THREADS = 320;
constexpr int G2_OFF_K = 2 * Q_BYTES, G2_OFF_V = G2_OFF_K + 2 * K_BYTES, G2_OFF_P = G2_OFF_V + 2 * V_BYTES;
constexpr int G2_OFF_BAR = G2_OFF_P + 4 * P_BYTES, G2_SMEM_BYTES = G2_OFF_BAR + 256 + 1024;   // P is double-buffered per group
constexpr int G2_TM_S = 0 /* +128 g */, G2_TM_O = 256 /* +64 g */;

template <int POLYMASK>      // bit c & 7 set: that 8-element chunk takes its exp2 on the FMA pipe (exp2_poly2), else on the MUFU
__global__ void __launch_bounds__(G2_THREADS, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_vt, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024 - (raw & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_OFF_BAR);
  uint64_t* q_full = bars;            // [1]
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;        // [2]
  uint64_t* v_empty = bars + 7;       // [2]
  uint64_t* s_full = bars + 9;        // [2] per group
  uint64_t* s_empty = bars + 11;      // [2]
  uint64_t* p_full = bars + 13;       // [2]
  uint64_t* p_empty = bars + 15;      // [2 groups][2 buffers]: PV of the tile that used this P buffer has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int T = p.L / BKV;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_vt)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      // softmax -> MMA barriers take ONE arrival per warp (lane 0 after __syncwarp): 128 per-thread arrivals on one shared-memory
      // word serialise
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[2 * i], 1); mbar_init(&p_empty[2 * i + 1], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * Q_BYTES);
      tma_load_4d(&map_q, q_full, smem, 0, qt * 2 * BQ, h, n);
      tma_load_4d(&map_q, q_full, smem + Q_BYTES, 0, qt * 2 * BQ + BQ, h, n);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], K_BYTES);
        tma_load_4d(&map_k, &k_full[st], smem + G2_OFF_K + st * K_BYTES, 0, j * BKV, h, n);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], V_BYTES);
        uint8_t* sv = smem + G2_OFF_V + st * V_BYTES;
        tma_load_3d(&map_vt, &v_full[st], sv, j * BKV, h * p.D, n);
        tma_load_3d(&map_vt, &v_full[st], sv + VBOX_BYTES, j * BKV + 64, h * p.D, n);
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BKV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const int kq = (p.D + 15) >> 4;      // k-steps of QK^T that hold data: q / k columns D..63 are zero, a 16-column step of zeros is skipped (D = 40: 3 of 4)
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      auto issue_s = [&](int j) {           // S_g(j) = Q_g K_j^T for both groups
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_full[st], ph);
        const uint64_t k_desc = sw128_desc(smem_u32(smem + G2_OFF_K + st * K_BYTES));
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&s_empty[g], (uint32_t)((j & 1) ^ 1));
          tc_fence_after();
          const uint64_t q_desc = sw128_desc(smem_u32(smem + g * Q_BYTES));
#pragma unroll
          for (int kk = 0; kk < kq; ++kk)
            umma(tmem_base + G2_TM_S + g * 128, q_desc + 2 * kk, k_desc + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
          tc_commit(&s_full[g]);
        }
        tc_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&v_full[st], ph);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_full[g], (uint32_t)(j & 1));
          tc_fence_after();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint64_t a_desc = sw128_desc(smem_u32(smem + G2_OFF_P + (2 * g + (j & 1)) * P_BYTES + hh * PHALF_BYTES));
            const uint64_t b_desc = sw128_desc(smem_u32(smem + G2_OFF_V + st * V_BYTES + hh * VBOX_BYTES));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma(tmem_base + G2_TM_O + g * 64, a_desc + 2 * kk, b_desc + 2 * kk, idesc_o, (j > 0 || hh > 0 || kk > 0) ? 1u : 0u);
          }
          tc_commit(&p_empty[2 * g + (j & 1)]);
        }
        tc_commit(&v_empty[st]);
      }
    }
  } else {
    // ================================================================== softmax / correction / epilogue (group g = warps 2-5 | 6-9)
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale_log2e;
    float m_used = -INFINITY, l = 0.f;
    const uint32_t s_addr = tmem_base + lane_base + G2_TM_S + g * 128;
    const uint32_t o_addr = tmem_base + lane_base + G2_TM_O + g * 64;
    for (int j = 0; j < T; ++j) {
      const uint32_t par = (uint32_t)(j & 1);
      mbar_wait(&s_full[g], par);
      tc_fence_after();
      uint32_t sr[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(s_addr + c * 32, sr + c * 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[g]);                // S_g may be overwritten by tile j+1
      float mt = __uint_as_float(sr[0]);
#pragma unroll
      for (int i = 1; i < 128; ++i) mt = fmaxf(mt, __uint_as_float(sr[i]));
      float factor = 1.0f;
      bool need = false;
      if (j == 0) {
        m_used = mt;
      } else if ((mt - m_used) * sl2 > 8.0f) {
        factor = ex2((m_used - mt) * sl2);
        m_used = mt;
        need = true;
      }
      l *= factor;
      // x = s * scale*log2e - m  in (-inf, 8]; clamped at -126 so that the FMA-pipe exp2 can patch the exponent with an integer add
      const float nbf = -m_used * sl2;
      const f32x2 sl2p = pk2(sl2, sl2), nbp = pk2(nbf, nbf);
      // P buffer j & 1 was last read by PV_g(j-2): with two buffers the exponentials of tile j overlap PV_g(j-1) instead of waiting
      mbar_wait(&p_empty[2 * g + (j & 1)], (uint32_t)(((j >> 1) & 1) ^ 1));
      uint8_t* prow = smem + G2_OFF_P + (2 * g + (j & 1)) * P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
      f32x2 sum2 = pk2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        uint32_t pw[4];
        const bool poly = ((POLYMASK >> (c & 7)) & 1) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const f32x2 x = fma2(pk2(__uint_as_float(sr[c * 8 + 2 * i]), __uint_as_float(sr[c * 8 + 2 * i + 1])), sl2p, nbp);
          f32x2 e;
          if (POLYMASK & 0x100) {              // DIAGNOSTIC build (FYC_ATTN_DBG): no exponentials - wrong results, times the rest of the pipeline
            e = x;
          } else if (poly) {
            float x0, x1; upk2(x, x0, x1);
            e = exp2_poly2(pk2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f)));
          } else {
            float x0, x1; upk2(x, x0, x1);
            e = pk2(ex2(x0), ex2(x1));
          }
          sum2 = add2(sum2, e);
          float e0, e1; upk2(e, e0, e1);
          __nv_bfloat162 t = __floats2bfloat162_rn(e0, e1);
          pw[i] = *reinterpret_cast<uint32_t*>(&t);
        }
        const int half = c >> 3, cc = c & 7;
        if (!(POLYMASK & 0x200) || c == 0)     // DIAGNOSTIC 0x200: one of the sixteen P stores only
          *reinterpret_cast<uint4*>(prow + half * PHALF_BYTES + ((cc ^ (r & 7)) << 4)) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
      }
      { float a0, a1; upk2(sum2, a0, a1); l += a0 + a1; }
      if (__any_sync(0xffffffffu, need)) {
        // O_g must be quiescent: PV_g(j-1), the last MMA that accumulates into it before PV_g(j) (which waits for this warp), retired
        mbar_wait(&p_empty[2 * g + ((j - 1) & 1)], (uint32_t)(((j - 1) >> 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < DV / 16; ++c) {
          uint32_t orr[16];
          tmem_ld16(o_addr + c * 16, orr);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * factor);
          tmem_st16(o_addr + c * 16, orr);
        }
        tmem_wait_st();
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
    }
    mbar_wait(&p_empty[2 * g + ((T - 1) & 1)], (uint32_t)(((T - 1) >> 1) & 1));
    tc_fence_after();
    const float inv = p.out_alpha / l;
    bf16* orow = p.out + (int64_t)n * p.bso + ((int64_t)qt * 2 * BQ + g * BQ + r) * p.ldo + h * p.D;
#pragma unroll
    for (int c = 0; c < DV / 16; ++c) {
      uint32_t orr[16];
      tmem_ld16(o_addr + c * 16, orr);
      tmem_wait_ld();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(orr[i]) * inv;
      if (c * 16 + 8 <= p.D) Vec8<bf16>::store(orow + c * 16, f);
      if (c * 16 + 16 <= p.D) Vec8<bf16>::store(orow + c * 16 + 8, f + 8);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Generalised ping-pong kernel with 64-key tiles: G query tiles (groups of 128 rows, one softmax warpgroup each) per CTA, head dims that
// span KA 64-column swizzle atoms, PV accumulators DVN columns wide.  Two instantiations:
//   <G = 2, KA = 2, DVN = 80>  head dim 80 (level-1 self-attention, 1024 tokens): q / k heads UNPADDED in the fused [q | k | v] buffer - the
//       head's second atom (columns 64..127 of its 160-byte-strided row) holds 16 of its own columns and 48 of the next head's, and QK^T
//       issues only the k-steps that hold data (5 x 16 columns), so the foreign columns are never multiplied.  Replaces the legacy
//       mma.sync kernel for this shape (214 TFLOP/s, profiles/round1_probe_c_shapes.txt).
//   <G = 4, KA = 1, DVN = 48>  head dim 40 with four softmax warps per scheduler - an EXPERIMENT that measured slower than
//       attention_tc2 (1.70 vs 1.43 ms at 32 x 8 x 4096, profiles/round2_attention.md): opt-in, FYC_ATTN_G4=1.
// TMEM: S_g G x 64 columns, O_g at 256 + g x (64 | 128).  Shared memory: Q G x KA x 16 KB, P G x 16 KB (one buffer per group: only the 16-byte
// stores of tile j + 1 wait for PV_g(j), the exponentials themselves do not), K / V^T rings of GK_STAGES x (KA x 8 KB + DVN x 128 B) -
// the 64-key tiles turn over in well under the TMA latency, so the ring is four deep.
constexpr int GK_BKV = 64, GK_STAGES = 4;
constexpr int GK_P_BYTES = PHALF_BYTES;                          // 16 KB: 128 rows x 64 keys
template <int G, int KA, int DVN> struct GkCfg {
  static constexpr int THREADS = 64 + G * 128;
  static constexpr int Q_TILE = KA * Q_BYTES;                     // per group
  static constexpr int K_STAGE = KA * GK_BKV * 128;               // KA atoms of 64 keys x 128 B
  static constexpr int V_STAGE = DVN * 128;                       // DVN rows (d) x 64 keys
  static constexpr int OFF_K = G * Q_TILE, OFF_V = OFF_K + GK_STAGES * K_STAGE, OFF_P = OFF_V + GK_STAGES * V_STAGE;
  static constexpr int OFF_BAR = OFF_P + G * GK_P_BYTES, SMEM = OFF_BAR + 512 + 1024;
  static_assert(V_STAGE % 1024 == 0 && K_STAGE % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte aligned bases");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static constexpr int O_STRIDE = DVN <= 64 ? 64 : 128;         // TMEM columns between the groups' O accumulators
  static_assert(G * 64 <= 256 && 256 + G * O_STRIDE <= 512, "TMEM column budget");
};

template <int G, int KA, int DVN, int POLYMASK>
__global__ void __launch_bounds__(GkCfg<G, KA, DVN>::THREADS, 1)
attention_tcg_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_vt, const AttnTcParams p) {
  using Cfg = GkCfg<G, KA, DVN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024 - (raw & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars;                          // [1]
  uint64_t* k_full = bars + 1;                      // [GK_STAGES]
  uint64_t* k_empty = k_full + GK_STAGES;
  uint64_t* v_full = k_empty + GK_STAGES;
  uint64_t* v_empty = v_full + GK_STAGES;
  uint64_t* s_full = v_empty + GK_STAGES;           // [G]  S_g(j) is in TMEM
  uint64_t* s_empty = s_full + G;                   // [G]  S_g has been read into registers
  uint64_t* p_full = s_empty + G;                   // [G]  P_g(j) is in shared memory (and O_g rescaled if needed)
  uint64_t* p_empty = p_full + G;                   // [G]  PV_g(j) has retired: P_g may be overwritten, O_g is quiescent
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + G);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int T = p.L / GK_BKV;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_vt)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < GK_STAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < G; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, G * Cfg::Q_TILE);
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int a = 0; a < KA; ++a)
          tma_load_4d(&map_q, q_full, smem + g * Cfg::Q_TILE + a * Q_BYTES, a * 64, qt * G * BQ + g * BQ, h, n);
      int st = 0; uint32_t ph = 0;
      for (int j = 0; j < T; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], Cfg::K_STAGE);
#pragma unroll
        for (int a = 0; a < KA; ++a)
          tma_load_4d(&map_k, &k_full[st], smem + Cfg::OFF_K + st * Cfg::K_STAGE + a * (GK_BKV * 128), a * 64, j * GK_BKV, h, n);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], Cfg::V_STAGE);
        tma_load_3d(&map_vt, &v_full[st], smem + Cfg::OFF_V + st * Cfg::V_STAGE, j * GK_BKV, h * p.D, n);
        if (++st == GK_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GK_BKV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DVN >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const int kq = (p.D + 15) >> 4;      // k-steps of QK^T that hold this head's data; later columns of the last atom are never read
      auto issue_s = [&](int j) {           // S_g(j) = Q_g K_j^T for all groups
        const int st = j % GK_STAGES; const uint32_t ph = (uint32_t)((j / GK_STAGES) & 1);
        mbar_wait(&k_full[st], ph);
        const uint32_t k_base = smem_u32(smem + Cfg::OFF_K + st * Cfg::K_STAGE);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          mbar_wait(&s_empty[g], (uint32_t)((j & 1) ^ 1));
          tc_fence_after();
          const uint32_t q_base = smem_u32(smem + g * Cfg::Q_TILE);
          for (int kk = 0; kk < kq; ++kk) {
            const int atom = kk >> 2, within = kk & 3;
            umma(tmem_base + g * 64, sw128_desc(q_base + atom * Q_BYTES) + 2 * within, sw128_desc(k_base + atom * (GK_BKV * 128)) + 2 * within,
                 idesc_s, kk > 0 ? 1u : 0u);
          }
          tc_commit(&s_full[g]);
        }
        tc_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        const int st = j % GK_STAGES; const uint32_t ph = (uint32_t)((j / GK_STAGES) & 1);
        mbar_wait(&v_full[st], ph);
        const uint64_t b_desc = sw128_desc(smem_u32(smem + Cfg::OFF_V + st * Cfg::V_STAGE));
#pragma unroll
        for (int g = 0; g < G; ++g) {
          mbar_wait(&p_full[g], (uint32_t)(j & 1));
          tc_fence_after();
          const uint64_t a_desc = sw128_desc(smem_u32(smem + Cfg::OFF_P + g * GK_P_BYTES));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma(tmem_base + 256 + g * Cfg::O_STRIDE, a_desc + 2 * kk, b_desc + 2 * kk, idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          tc_commit(&p_empty[g]);
        }
        tc_commit(&v_empty[st]);
      }
    }
  } else {
    // ================================================================== softmax / correction / epilogue (group g = warps 2 + 4g .. 5 + 4g)
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale_log2e;
    float m_used = -INFINITY, l = 0.f;
    const uint32_t s_addr = tmem_base + lane_base + g * 64;
    const uint32_t o_addr = tmem_base + lane_base + 256 + g * Cfg::O_STRIDE;
    uint8_t* const prow = smem + Cfg::OFF_P + g * GK_P_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
    for (int j = 0; j < T; ++j) {
      const uint32_t par = (uint32_t)(j & 1);
      mbar_wait(&s_full[g], par);
      tc_fence_after();
      uint32_t sr[64];
      tmem_ld32(s_addr, sr);
      tmem_ld32(s_addr + 32, sr + 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[g]);                // S_g may be overwritten by tile j + 1
      float mt = __uint_as_float(sr[0]);
#pragma unroll
      for (int i = 1; i < 64; ++i) mt = fmaxf(mt, __uint_as_float(sr[i]));
      float factor = 1.0f;
      bool need = false;
      if (j == 0) {
        m_used = mt;
      } else if ((mt - m_used) * sl2 > 8.0f) {                // lazy rescale: O is touched only when the running max grows by > 2^8
        factor = ex2((m_used - mt) * sl2);
        m_used = mt;
        need = true;
      }
      l *= factor;
      const float nbf = -m_used * sl2;
      const f32x2 sl2p = pk2(sl2, sl2), nbp = pk2(nbf, nbf);
      f32x2 sum2 = pk2(0.f, 0.f);
      uint32_t pw[32];                                         // the row's 64 probabilities as bf16 pairs, stored once P_g is free
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const bool poly = ((POLYMASK >> c) & 1) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const f32x2 x = fma2(pk2(__uint_as_float(sr[c * 8 + 2 * i]), __uint_as_float(sr[c * 8 + 2 * i + 1])), sl2p, nbp);
          f32x2 e;
          float x0, x1; upk2(x, x0, x1);
          if (poly) e = exp2_poly2(pk2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f)));
          else e = pk2(ex2(x0), ex2(x1));
          sum2 = add2(sum2, e);
          float e0, e1; upk2(e, e0, e1);
          __nv_bfloat162 t = __floats2bfloat162_rn(e0, e1);
          pw[c * 4 + i] = *reinterpret_cast<uint32_t*>(&t);
        }
      }
      { float a0, a1; upk2(sum2, a0, a1); l += a0 + a1; }
      // PV_g(j - 1) must have retired before P_g is overwritten and before O_g may be rescaled (one wait serves both)
      if (j > 0) mbar_wait(&p_empty[g], (uint32_t)((j - 1) & 1));
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = make_uint4(pw[c * 4], pw[c * 4 + 1], pw[c * 4 + 2], pw[c * 4 + 3]);
      if (__any_sync(0xffffffffu, need)) {
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < DVN / 16; ++c) {
          uint32_t orr[16];
          tmem_ld16(o_addr + c * 16, orr);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * factor);
          tmem_st16(o_addr + c * 16, orr);
        }
        tmem_wait_st();
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
    }
    mbar_wait(&p_empty[g], (uint32_t)((T - 1) & 1));
    tc_fence_after();
    const float inv = p.out_alpha / l;
    bf16* orow = p.out + (int64_t)n * p.bso + ((int64_t)qt * G * BQ + g * BQ + r) * p.ldo + h * p.D;
#pragma unroll
    for (int c = 0; c < DVN / 16; ++c) {
      uint32_t orr[16];
      tmem_ld16(o_addr + c * 16, orr);
      tmem_wait_ld();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(orr[i]) * inv;
      if (c * 16 + 8 <= p.D) Vec8<bf16>::store(orow + c * 16, f);
      if (c * 16 + 16 <= p.D) Vec8<bf16>::store(orow + c * 16 + 8, f + 8);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Cross-attention on tcgen05: a SHORT, step-invariant context (the 77 text tokens, padded to 80 keys; optionally the IP-Adapter's
// 4 / 16 image tokens, padded to 16) against long query sequences.  Replaces the mma.sync `attention_mma_shortk_kernel` for head dims
// 40 and 80 (CrossAttention._attention diffusers/models/attention.py:649-678 for attn2; IPCrossAttention.forward
// animatediff/models/attention.py:92-120 == IPAttnProcessor.__call__ ip_adapter/attention_processor.py:137-168).
//
// One CTA = one (image, head): K_t, V_t^T (and K_i, V_i^T) are loaded ONCE and stay in shared memory; two softmax warpgroups ping-pong
// over the CTA's query tiles (128 rows each).  Per tile:  S_t = Q K_t^T (N = 80) and S_i = Q K_i^T (N = 16) into TMEM -> one row per
// thread: two INDEPENDENT softmaxes, normalised in registers (the whole context is one tile: no online rescaling) and pre-scaled by
// out_alpha / l_t and alpha2 / l_i -> P_t, P_i (bf16) to shared memory -> O = P_t V_t + P_i V_i accumulated in ONE TMEM accumulator ->
// bf16 out, written once.  Padding keys (77..79, T..15) are masked to -inf before the softmax.
// Operands (prepared once per clip by the host, unet.py::prepare_context): q UNPADDED [NB, Lq, heads * D] (head h at columns D h): the
// 64-column TMA box of a head reads 24 foreign columns for D = 40, of which k-step 2 multiplies columns 40..47 - against ZERO columns
// of K, which the host pads per head to 64 (D = 80: two atoms, 5 k-steps, no padding, as in the self-attention kernel above);
// k [NBc, 80, heads * DKP] (DKP = 64 | 80), vt [NBc, heads * D, 80]; k2 [NBc, 16, heads * DKP], vt2 [NBc, heads * D, 16].
template <int KA, int DVN> struct CxCfg {
  static constexpr int LKT = 80, LKI = 16;                        // padded key counts (text, image)
  static constexpr int THREADS = 64 + 2 * 128;
  static constexpr int KT_BYTES = KA * LKT * 128;                 // K_t: KA atoms of 80 keys x 128 B
  static constexpr int VT_BYTES = 2 * DVN * 128;                  // V_t^T: two 64-key atoms (keys 0..63 | 64..79 + zero fill) of DVN rows
  static constexpr int KI_BYTES = KA * LKI * 128;
  static constexpr int VI_BYTES = DVN * 128;
  static constexpr int Q_TILE = KA * Q_BYTES;                     // per group
  static constexpr int P_TILE = 3 * PHALF_BYTES;                  // per group: P_t atoms 0, 1 and P_i
  static constexpr int OFF_KT = 0, OFF_VT = OFF_KT + KT_BYTES, OFF_KI = OFF_VT + VT_BYTES, OFF_VI = OFF_KI + KI_BYTES;
  static constexpr int OFF_Q = OFF_VI + VI_BYTES, OFF_P = OFF_Q + 2 * Q_TILE, OFF_BAR = OFF_P + 2 * P_TILE, SMEM = OFF_BAR + 256 + 1024;
  static constexpr int TM_ST = 0, TM_SI = 96, TM_O = 128, TM_GROUP = 256;      // TMEM columns inside a group's 256
  static_assert(KT_BYTES % 1024 == 0 && VT_BYTES % 1024 == 0 && KI_BYTES % 1024 == 0 && VI_BYTES % 1024 == 0, "SWIZZLE_128B tile bases");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(TM_O + DVN <= TM_GROUP, "TMEM column budget");
};

struct AttnCxParams {
  bf16* out; int64_t ldo, bso;
  int Lq, heads, D, Lk, Lk2;        // Lk <= 80 real text keys, Lk2 = 0 (no second context) or <= 16 real image keys
  int dkp;                          // column stride between the heads of k / k2 (64 for D = 40: zero-padded heads; 80 for D = 80)
  int kv_div;                       // images per context (= F)
  float scale_log2e, out_alpha, alpha2;
};

template <int KA, int DVN, int POLYMASK>
__global__ void __launch_bounds__(CxCfg<KA, DVN>::THREADS, 1)
attention_cx_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_vt,
                    const __grid_constant__ CUtensorMap map_k2, const __grid_constant__ CUtensorMap map_vt2, const AttnCxParams p) {
  using Cfg = CxCfg<KA, DVN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024 - (raw & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* kv_full = bars;              // [1]
  uint64_t* q_full = bars + 1;           // [2] per group: the group's Q tile has landed
  uint64_t* q_empty = bars + 3;          // [2] QK^T of the tile has retired: the Q slot may be refilled
  uint64_t* s_full = bars + 5;           // [2] S_t, S_i are in TMEM
  uint64_t* s_empty = bars + 7;          // [2] ... and have been read into registers
  uint64_t* p_full = bars + 9;           // [2] P_t, P_i are in shared memory and O_g has been drained
  uint64_t* o_full = bars + 11;          // [2] PV has retired: O_g is complete, P_g may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, n = blockIdx.z;
  const int nc = n / p.kv_div;
  const int nqt = (p.Lq + BQ - 1) / BQ;
  const bool ip = p.Lk2 > 0;
  // this CTA's query tiles: blockIdx.x, + gridDim.x, ...; group g takes every other one of them
  auto tile_of = [&](int g, int it) { return (int)blockIdx.x + (2 * it + g) * (int)gridDim.x; };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_vt)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&q_full[g], 1); mbar_init(&q_empty[g], 1); mbar_init(&s_full[g], 1); mbar_init(&s_empty[g], 4);
      mbar_init(&p_full[g], 4); mbar_init(&o_full[g], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      const uint32_t kv_bytes = Cfg::KT_BYTES + Cfg::VT_BYTES + (ip ? Cfg::KI_BYTES + Cfg::VI_BYTES : 0);
      mbar_expect_tx(kv_full, kv_bytes);
#pragma unroll
      for (int a = 0; a < KA; ++a) tma_load_3d(&map_k, kv_full, smem + Cfg::OFF_KT + a * (Cfg::LKT * 128), h * p.dkp + a * 64, 0, nc);
      tma_load_3d(&map_vt, kv_full, smem + Cfg::OFF_VT, 0, h * p.D, nc);
      tma_load_3d(&map_vt, kv_full, smem + Cfg::OFF_VT + DVN * 128, 64, h * p.D, nc);
      if (ip) {
#pragma unroll
        for (int a = 0; a < KA; ++a) tma_load_3d(&map_k2, kv_full, smem + Cfg::OFF_KI + a * (Cfg::LKI * 128), h * p.dkp + a * 64, 0, nc);
        tma_load_3d(&map_vt2, kv_full, smem + Cfg::OFF_VI, 0, h * p.D, nc);
      }
      for (int it = 0;; ++it) {
        bool any = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int t = tile_of(g, it);
          if (t >= nqt) continue;
          any = true;
          mbar_wait(&q_empty[g], (uint32_t)((it & 1) ^ 1));
          mbar_expect_tx(&q_full[g], Cfg::Q_TILE);
#pragma unroll
          for (int a = 0; a < KA; ++a) tma_load_3d(&map_q, &q_full[g], smem + Cfg::OFF_Q + g * Cfg::Q_TILE + a * Q_BYTES, h * p.D + a * 64, t * BQ, n);
        }
        if (!any) break;
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc_st = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cfg::LKT >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint32_t idesc_si = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cfg::LKI >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DVN >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
      const int kq = (p.D + 15) >> 4;
      const uint32_t kt_base = smem_u32(smem + Cfg::OFF_KT), ki_base = smem_u32(smem + Cfg::OFF_KI);
      mbar_wait(kv_full, 0);
      auto issue_s = [&](int g, int it) {
        mbar_wait(&q_full[g], (uint32_t)(it & 1));
        mbar_wait(&s_empty[g], (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
        const uint32_t q_base = smem_u32(smem + Cfg::OFF_Q + g * Cfg::Q_TILE);
        const uint32_t tm = tmem_base + g * Cfg::TM_GROUP;
        for (int kk = 0; kk < kq; ++kk) {
          const int atom = kk >> 2, within = kk & 3;
          umma(tm + Cfg::TM_ST, sw128_desc(q_base + atom * Q_BYTES) + 2 * within, sw128_desc(kt_base + atom * (Cfg::LKT * 128)) + 2 * within, idesc_st, kk > 0 ? 1u : 0u);
        }
        if (ip)
          for (int kk = 0; kk < kq; ++kk) {
            const int atom = kk >> 2, within = kk & 3;
            umma(tm + Cfg::TM_SI, sw128_desc(q_base + atom * Q_BYTES) + 2 * within, sw128_desc(ki_base + atom * (Cfg::LKI * 128)) + 2 * within, idesc_si, kk > 0 ? 1u : 0u);
          }
        tc_commit(&s_full[g]);
        tc_commit(&q_empty[g]);
      };
      auto issue_o = [&](int g, int it) {
        mbar_wait(&p_full[g], (uint32_t)(it & 1));
        tc_fence_after();
        const uint32_t p_base = smem_u32(smem + Cfg::OFF_P + g * Cfg::P_TILE);
        const uint32_t tm = tmem_base + g * Cfg::TM_GROUP + Cfg::TM_O;
        const uint32_t vt_base = smem_u32(smem + Cfg::OFF_VT);
#pragma unroll
        for (int kk = 0; kk < Cfg::LKT / 16; ++kk) {       // 80 keys: four k-steps in atom 0, one in atom 1
          const int atom = kk >> 2, within = kk & 3;
          umma(tm, sw128_desc(p_base + atom * PHALF_BYTES) + 2 * within, sw128_desc(vt_base + atom * (DVN * 128)) + 2 * within, idesc_o, kk > 0 ? 1u : 0u);
        }
        if (ip) umma(tm, sw128_desc(p_base + 2 * PHALF_BYTES), sw128_desc(smem_u32(smem + Cfg::OFF_VI)), idesc_o, 1u);
        tc_commit(&o_full[g]);
      };
      // order: S(g0, 0), S(g1, 0), then per iteration O(g, it) followed by S(g, it + 1) - the other group's softmax overlaps
      for (int g = 0; g < 2; ++g) if (tile_of(g, 0) < nqt) issue_s(g, 0);
      for (int it = 0;; ++it) {
        bool any = false;
        for (int g = 0; g < 2; ++g) {
          if (tile_of(g, it) >= nqt) continue;
          any = true;
          issue_o(g, it);
          if (tile_of(g, it + 1) < nqt) issue_s(g, it + 1);
        }
        if (!any) break;
      }
    }
  } else {
    // ================================================================== softmax + epilogue (group g = warps 2-5 | 6-9)
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale_log2e;
    const uint32_t tm = tmem_base + lane_base + g * Cfg::TM_GROUP;
    uint8_t* const prow = smem + Cfg::OFF_P + g * Cfg::P_TILE + (r >> 3) * 1024 + (r & 7) * 128;
    for (int it = 0;; ++it) {
      const int t = tile_of(g, it);
      if (t >= nqt) break;
      const uint32_t par = (uint32_t)(it & 1);
      mbar_wait(&s_full[g], par);
      tc_fence_after();
      uint32_t st[80], si[16];
      tmem_ld32(tm + Cfg::TM_ST, st);
      tmem_ld32(tm + Cfg::TM_ST + 32, st + 32);
      tmem_ld16(tm + Cfg::TM_ST + 64, st + 64);
      if (ip) tmem_ld16(tm + Cfg::TM_SI, si);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[g]);
      // ---- text softmax over Lk keys (padding keys are TMA-zero rows of K: score 0 -> masked here)
      float mt = -INFINITY;
#pragma unroll
      for (int i = 0; i < 80; ++i) { if (i >= p.Lk) st[i] = 0xff800000u; mt = fmaxf(mt, __uint_as_float(st[i])); }
      const float nb = -mt * sl2;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 80; ++i) { const float e = ex2(fmaf(__uint_as_float(st[i]), sl2, nb)); sum += e; st[i] = __float_as_uint(e); }
      const float wt = p.out_alpha / sum;
      float mi = -INFINITY, wi = 0.f;
      if (ip) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { if (i >= p.Lk2) si[i] = 0xff800000u; mi = fmaxf(mi, __uint_as_float(si[i])); }
        const float nbi = -mi * sl2;
        float sumi = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float e = ex2(fmaf(__uint_as_float(si[i]), sl2, nbi)); sumi += e; si[i] = __float_as_uint(e); }
        wi = p.alpha2 / sumi;
      }
      // ---- P (normalised, weighted) -> shared memory; the previous tile's PV must have retired (it also means O was ... see below)
      if (it > 0) mbar_wait(&o_full[g], (uint32_t)((it - 1) & 1));      // (already passed in the epilogue of tile it - 1: kept for clarity)
#pragma unroll
      for (int c = 0; c < 10; ++c) {                   // 80 keys = 10 chunks of 8: chunks 0-7 -> atom 0, 8-9 -> atom 1
        uint32_t pw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(st[c * 8 + 2 * i]) * wt, __uint_as_float(st[c * 8 + 2 * i + 1]) * wt);
          pw[i] = *reinterpret_cast<uint32_t*>(&v);
        }
        const int atom = c >> 3, cc = c & 7;
        *reinterpret_cast<uint4*>(prow + atom * PHALF_BYTES + ((cc ^ (r & 7)) << 4)) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
      }
      if (ip) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t pw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(si[c * 8 + 2 * i]) * wi, __uint_as_float(si[c * 8 + 2 * i + 1]) * wi);
            pw[i] = *reinterpret_cast<uint32_t*>(&v);
          }
          *reinterpret_cast<uint4*>(prow + 2 * PHALF_BYTES + ((c ^ (r & 7)) << 4)) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      // ---- epilogue of THIS tile: O = P_t V_t + P_i V_i is already normalised and weighted
      mbar_wait(&o_full[g], par);
      tc_fence_after();
      const int64_t row = (int64_t)t * BQ + r;
      bf16* orow = p.out + (int64_t)n * p.bso + row * p.ldo + h * p.D;
#pragma unroll
      for (int c = 0; c < DVN / 16; ++c) {
        uint32_t orr[16];
        tmem_ld16(tm + Cfg::TM_O + c * 16, orr);
        tmem_wait_ld();
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(orr[i]);
        if (row < p.Lq) {
          if (c * 16 + 8 <= p.D) Vec8<bf16>::store(orow + c * 16, f);
          if (c * 16 + 16 <= p.D) Vec8<bf16>::store(orow + c * 16 + 8, f + 8);
        }
      }
      tc_fence_before();      // the O reads are ordered before the next p_full arrival (PV of the next tile overwrites O only after it)
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// [NB, L, ld] (columns col0 .. col0+C) -> [NB, C, L]   (V -> V^T so that keys are the contiguous, K-major dimension of PV)
__global__ void __launch_bounds__(256) transpose_tokens_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int L, int C,
                                                               int64_t ld, int64_t col0) {
  __shared__ bf16 tile[64][66];
  const int n = blockIdx.z, t0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const bf16* src = in + (int64_t)n * L * ld + col0;
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {        // 64 tokens x 32 channel pairs
    int t = i >> 5, cp = (i & 31) * 2;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
    if (t0 + t < L && c0 + cp < C) v = *reinterpret_cast<const __nv_bfloat162*>(src + (int64_t)(t0 + t) * ld + c0 + cp);
    tile[t][cp] = v.x; tile[t][cp + 1] = v.y;
  }
  __syncthreads();
  bf16* dst = out + (int64_t)n * C * L;
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {        // 64 channels x 32 token pairs
    int c = i >> 5, tp = (i & 31) * 2;
    if (c0 + c < C && t0 + tp < L) {
      __nv_bfloat162 v;
      v.x = tile[tp][c]; v.y = tile[tp + 1][c];
      *reinterpret_cast<__nv_bfloat162*>(dst + (int64_t)(c0 + c) * L + t0 + tp) = v;
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    else
      (void)cudaGetLastError();
  }
  return fn;
}
int32_t make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  FYC_CHECK(fn != nullptr, "attention(tcgen05): cuTensorMapEncodeTiled unavailable");
  cuuint64_t gd[5], gs[4]; cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FYC_CHECK(r == CUDA_SUCCESS, "attention(tcgen05): cuTensorMapEncodeTiled failed (%d)", (int)r);
  return FYC_OK;
}

}  // namespace

extern "C" int32_t fyc_transpose_tokens(const void* in, void* out, int64_t NB, int64_t L, int64_t C, int64_t ld, int64_t col0,
                                        void* stream) {
  FYC_CHECK(C % 2 == 0 && L % 2 == 0 && ld % 2 == 0 && col0 % 2 == 0 && NB < 65536, "transpose_tokens: even sizes required");
  dim3 grid((unsigned)ceil_div64(L, 64), (unsigned)ceil_div64(C, 64), (unsigned)NB);
  transpose_tokens_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)in, (bf16*)out, (int)L, (int)C, ld, col0);
  FYC_LAUNCH_CHECK();
  return FYC_OK;
}

// qk: [NB, L, ldqk] bf16 with q head h at columns [q_col0 + 64h, +64) and k head h at [k_col0 + 64h, +64) (cols D..63 zero);
// vt: [NB, heads*D, L]; out: [NB, L, ldo] (head h at columns [h*D, (h+1)*D)).
extern "C" int32_t fyc_self_attention_tc(const void* qk, int64_t ldqk, int64_t q_col0, int64_t k_col0, const void* vt, void* out,
                                         int64_t ldo, int64_t NB, int64_t heads, int64_t L, int64_t D, float scale, void* stream) {
  FYC_CHECK(D == 40, "self_attention_tc: built for head dim 40 (got %lld)", (long long)D);
  FYC_CHECK(L % 128 == 0 && L >= 128, "self_attention_tc: sequence length %lld must be a multiple of 128", (long long)L);
  FYC_CHECK(ldqk % 8 == 0 && q_col0 % 8 == 0 && k_col0 % 8 == 0 && ldo % 8 == 0, "self_attention_tc: 16-byte alignment");
  FYC_CHECK((((uintptr_t)qk | (uintptr_t)vt | (uintptr_t)out) & 15) == 0, "self_attention_tc: pointers must be 16-byte aligned");
  FYC_CHECK(NB < 65536 && heads < 65536, "self_attention_tc: grid too large");
  CUtensorMap mq, mk, mv;
  {
    uint64_t dims[4] = {64, (uint64_t)L, (uint64_t)heads, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)ldqk * 2, 128, (uint64_t)L * ldqk * 2};
    uint32_t box[4] = {64, 128, 1, 1};
    int32_t rc = make_map(&mq, (const bf16*)qk + q_col0, 4, dims, str, box);
    if (rc) return rc;
    rc = make_map(&mk, (const bf16*)qk + k_col0, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)L, (uint64_t)(heads * D), (uint64_t)NB};
    uint64_t str[2] = {(uint64_t)L * 2, (uint64_t)L * heads * D * 2};
    uint32_t box[3] = {64, (uint32_t)DV, 1};
    int32_t rc = make_map(&mv, vt, 3, dims, str, box);
    if (rc) return rc;
  }
  static bool attr = false;
  if (!attr) {
    FYC_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  AttnTcParams p;
  p.out = (bf16*)out; p.ldo = ldo; p.bso = L * ldo; p.L = (int)L; p.heads = (int)heads; p.D = (int)D;
  p.scale_log2e = scale * 1.4426950408889634f; p.out_alpha = 1.0f;
  const char* g4e = getenv("FYC_ATTN_G4");
  // Four query tiles per CTA with 64-key tiles: measured SLOWER than the two-tile kernel (1.70 vs 1.43 ms at 32 x 8 x 4096, round 2:
  // profiles/round2_attention.md - half of its warp samples wait for MMA completions, the 64-key tiles double the MMA instruction count and
  // the shared-memory operand reads per key); kept as an opt-in experiment, FYC_ATTN_G4=1.
  if (L % (4 * BQ) == 0 && g4e && g4e[0] == '1') {
    using Cfg = GkCfg<4, 1, DV>;
    CUtensorMap mk64;
    {
      uint64_t dims[4] = {64, (uint64_t)L, (uint64_t)heads, (uint64_t)NB};
      uint64_t str[3] = {(uint64_t)ldqk * 2, 128, (uint64_t)L * ldqk * 2};
      uint32_t box[4] = {64, (uint32_t)GK_BKV, 1, 1};
      int32_t rc = make_map(&mk64, (const bf16*)qk + k_col0, 4, dims, str, box);
      if (rc) return rc;
    }
    auto kern = attention_tcg_kernel<4, 1, DV, 0x52>;
    static bool attr4 = false;
    if (!attr4) { FYC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr4 = true; }
    dim3 grid4((unsigned)(L / (4 * BQ)), (unsigned)heads, (unsigned)NB);
    kern<<<grid4, Cfg::THREADS, Cfg::SMEM, (cudaStream_t)stream>>>(mq, mk64, mv, p);
    FYC_LAUNCH_CHECK();
    return FYC_OK;
  }
  if (L % (2 * BQ) == 0) {       // two query tiles per CTA (ping-pong softmax warpgroups)
    static bool attr2 = false;
    if (!attr2) {
      FYC_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<0x00>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
      FYC_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<0x02>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
      FYC_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<0x12>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
      FYC_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<0x52>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
      attr2 = true;
    }
    dim3 grid2((unsigned)(L / (2 * BQ)), (unsigned)heads, (unsigned)NB);
    // share of the exponentials moved from the MUFU to the FMA pipe, in eighths (FYC_ATTN_POLY=0..3 overrides; measured optimum below)
    const char* pe = getenv("FYC_ATTN_POLY");
    const int eighths = pe ? atoi(pe) : 3;          // 0: 1.574 ms, 1: 1.480, 2: 1.452, 3: 1.422 ms at 32 x 8 heads x 4096 tokens
    cudaStream_t s2 = (cudaStream_t)stream;
    const char* dbg = getenv("FYC_ATTN_DBG");      // diagnostics only (WRONG results): 1 = no exponentials, 2 = 1/16 of the P stores, 3 = both
    if (dbg && dbg[0] >= '1' && dbg[0] <= '3') {
      static bool attrd = false;
      if (!attrd) {
        FYC_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<0x152>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
        FYC_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<0x252>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
        FYC_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<0x352>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
        attrd = true;
      }
      if (dbg[0] == '1') attention_tc2_kernel<0x152><<<grid2, G2_THREADS, G2_SMEM_BYTES, s2>>>(mq, mk, mv, p);
      else if (dbg[0] == '2') attention_tc2_kernel<0x252><<<grid2, G2_THREADS, G2_SMEM_BYTES, s2>>>(mq, mk, mv, p);
      else attention_tc2_kernel<0x352><<<grid2, G2_THREADS, G2_SMEM_BYTES, s2>>>(mq, mk, mv, p);
      FYC_LAUNCH_CHECK();
      return FYC_OK;
    }
    if (eighths <= 0) attention_tc2_kernel<0x00><<<grid2, G2_THREADS, G2_SMEM_BYTES, s2>>>(mq, mk, mv, p);
    else if (eighths == 1) attention_tc2_kernel<0x02><<<grid2, G2_THREADS, G2_SMEM_BYTES, s2>>>(mq, mk, mv, p);
    else if (eighths == 2) attention_tc2_kernel<0x12><<<grid2, G2_THREADS, G2_SMEM_BYTES, s2>>>(mq, mk, mv, p);
    else attention_tc2_kernel<0x52><<<grid2, G2_THREADS, G2_SMEM_BYTES, s2>>>(mq, mk, mv, p);
    FYC_LAUNCH_CHECK();
    return FYC_OK;
  }
  dim3 grid((unsigned)(L / BQ), (unsigned)heads, (unsigned)NB);
  attention_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(mq, mk, mv, p);
  FYC_LAUNCH_CHECK();
  return FYC_OK;
}

// Head dim 80 (level-1 self-attention): qkv [NB, L, ldqkv] bf16 with q head h at columns [q_col0 + 80 h, +80), k at [k_col0 + 80 h, +80) -
// UNPADDED, the fused [q | k | v] projection as the GEMM wrote it (reads up to column k_col0 + 80 heads + 47: the buffer must hold at
// least 48 more columns after k's last head, which the v block provides); vt: [NB, heads * 80, L]; out: [NB, L, ldo].
extern "C" int32_t fyc_self_attention_tc_d80(const void* qkv, int64_t ldqkv, int64_t q_col0, int64_t k_col0, const void* vt, void* out,
                                             int64_t ldo, int64_t NB, int64_t heads, int64_t L, float scale, void* stream) {
  constexpr int D = 80;
  using Cfg = GkCfg<2, 2, D>;
  FYC_CHECK(L % (2 * BQ) == 0 && L >= 2 * BQ, "self_attention_tc_d80: sequence length %lld must be a multiple of 256", (long long)L);
  FYC_CHECK(ldqkv % 8 == 0 && q_col0 % 8 == 0 && k_col0 % 8 == 0 && ldo % 8 == 0, "self_attention_tc_d80: 16-byte alignment");
  FYC_CHECK((((uintptr_t)qkv | (uintptr_t)vt | (uintptr_t)out) & 15) == 0, "self_attention_tc_d80: pointers must be 16-byte aligned");
  FYC_CHECK(NB < 65536 && heads < 65536, "self_attention_tc_d80: grid too large");
  FYC_CHECK(q_col0 + heads * D + 48 <= ldqkv && k_col0 + heads * D + 48 <= ldqkv, "self_attention_tc_d80: the row must extend 48 columns past the last head");
  CUtensorMap mq, mk, mv;
  {
    // inner extent 128 columns per head although heads are 160 bytes apart: overlapping tensor-map dimensions are legal, and the second
    // atom's foreign columns are never multiplied (5 k-steps)
    uint64_t dims[4] = {128, (uint64_t)L, (uint64_t)heads, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)ldqkv * 2, (uint64_t)D * 2, (uint64_t)L * ldqkv * 2};
    uint32_t boxq[4] = {64, (uint32_t)BQ, 1, 1}, boxk[4] = {64, (uint32_t)GK_BKV, 1, 1};
    int32_t rc = make_map(&mq, (const bf16*)qkv + q_col0, 4, dims, str, boxq);
    if (rc) return rc;
    rc = make_map(&mk, (const bf16*)qkv + k_col0, 4, dims, str, boxk);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)L, (uint64_t)(heads * D), (uint64_t)NB};
    uint64_t str[2] = {(uint64_t)L * 2, (uint64_t)L * heads * D * 2};
    uint32_t box[3] = {64, (uint32_t)D, 1};
    int32_t rc = make_map(&mv, vt, 3, dims, str, box);
    if (rc) return rc;
  }
  AttnTcParams p;
  p.out = (bf16*)out; p.ldo = ldo; p.bso = L * ldo; p.L = (int)L; p.heads = (int)heads; p.D = D;
  p.scale_log2e = scale * 1.4426950408889634f; p.out_alpha = 1.0f;
  auto kern = attention_tcg_kernel<2, 2, D, 0x52>;
  static bool attr = false;
  if (!attr) { FYC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr = true; }
  dim3 grid((unsigned)(L / (2 * BQ)), (unsigned)heads, (unsigned)NB);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM, (cudaStream_t)stream>>>(mq, mk, mv, p);
  FYC_LAUNCH_CHECK();
  return FYC_OK;
}

// Cross-attention with a resident short context on tcgen05 (head dim 40 or 80).  q: [NB, Lq, ldq] bf16, head h at columns [q_col0 + D h, +D),
// UNPADDED, ldq >= heads * D;
// k: [NBc, 80, ldk] with head h at columns [DKP h, +D), DKP = 64 for D = 40 (columns D..63 ZERO) or 80 for D = 80, rows Lk..79 zero;
// vt: [NBc, heads * D, 80]; optional second context k2 [NBc, 16, ldk2], vt2 [NBc, heads * D, 16] (rows / columns Lk2..15 zero).
// out[n, i, h D + :] = out_alpha softmax_j<Lk(scale q k^T) v + alpha2 softmax_j<Lk2(scale q k2^T) v2, NBc = NB / kv_batch_div.
extern "C" int32_t fyc_cross_attention_tc(const void* q, int64_t ldq, int64_t q_col0, const void* k, int64_t ldk, const void* vt,
                                          const void* k2, int64_t ldk2, const void* vt2, void* out, int64_t ldo, int64_t NB, int64_t heads,
                                          int64_t Lq, int64_t D, int64_t Lk, int64_t Lk2, int64_t kv_batch_div, float scale, float out_alpha,
                                          float alpha2, void* stream) {
  FYC_CHECK(D == 40 || D == 80, "cross_attention_tc: head dim %lld (40 or 80)", (long long)D);
  FYC_CHECK(Lk >= 1 && Lk <= 80 && Lk2 >= 0 && Lk2 <= 16 && Lq >= 1 && kv_batch_div >= 1 && NB % kv_batch_div == 0, "cross_attention_tc: bad shape");
  FYC_CHECK((k2 != nullptr) == (Lk2 > 0) && (vt2 != nullptr) == (Lk2 > 0), "cross_attention_tc: second context needs k2, vt2 and Lk2 > 0");
  FYC_CHECK(ldq % 8 == 0 && q_col0 % 8 == 0 && ldk % 8 == 0 && ldo % 8 == 0 && (k2 == nullptr || ldk2 % 8 == 0), "cross_attention_tc: 16-byte alignment");
  FYC_CHECK((((uintptr_t)q | (uintptr_t)k | (uintptr_t)vt | (uintptr_t)out | (uintptr_t)k2 | (uintptr_t)vt2) & 15) == 0, "cross_attention_tc: pointers must be 16-byte aligned");
  FYC_CHECK(NB < 65536 && heads < 65536, "cross_attention_tc: grid too large");
  const int64_t NBc = NB / kv_batch_div;
  const int KA = D == 40 ? 1 : 2;
  const int64_t DKP = D == 40 ? 64 : 80;
  CUtensorMap mq, mk, mv, mk2, mv2;
  {
    // q and k as 3-D maps over the WHOLE head-packed row (box origin = the head's first column): columns past the row end - the tail
    // of the last head's 64-column atom - are out of bounds for TMA and zero-filled instead of read
    uint64_t dims[3] = {(uint64_t)(heads * D), (uint64_t)Lq, (uint64_t)NB};
    uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)Lq * ldq * 2};
    uint32_t box[3] = {64, (uint32_t)BQ, 1};
    int32_t rc = make_map(&mq, (const bf16*)q + q_col0, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)(heads * DKP), 80, (uint64_t)NBc};
    uint64_t str[2] = {(uint64_t)ldk * 2, (uint64_t)80 * ldk * 2};
    uint32_t box[3] = {64, 80, 1};
    int32_t rc = make_map(&mk, k, 3, dims, str, box);
    if (rc) return rc;
    uint64_t vd[3] = {80, (uint64_t)(heads * D), (uint64_t)NBc};
    uint64_t vs[2] = {160, (uint64_t)(80 * heads * D * 2)};
    uint32_t vb[3] = {64, (uint32_t)(D == 40 ? 48 : 80), 1};
    rc = make_map(&mv, vt, 3, vd, vs, vb);
    if (rc) return rc;
  }
  mk2 = mk; mv2 = mv;
  if (k2) {
    uint64_t dims[3] = {(uint64_t)(heads * DKP), 16, (uint64_t)NBc};
    uint64_t str[2] = {(uint64_t)ldk2 * 2, (uint64_t)16 * ldk2 * 2};
    uint32_t box[3] = {64, 16, 1};
    int32_t rc = make_map(&mk2, k2, 3, dims, str, box);
    if (rc) return rc;
    uint64_t vd[3] = {16, (uint64_t)(heads * D), (uint64_t)NBc};
    uint64_t vs[2] = {32, (uint64_t)(16 * heads * D * 2)};
    uint32_t vb[3] = {64, (uint32_t)(D == 40 ? 48 : 80), 1};
    rc = make_map(&mv2, vt2, 3, vd, vs, vb);
    if (rc) return rc;
  }
  AttnCxParams p;
  p.out = (bf16*)out; p.ldo = ldo; p.bso = Lq * ldo; p.Lq = (int)Lq; p.heads = (int)heads; p.D = (int)D; p.Lk = (int)Lk; p.Lk2 = (int)Lk2;
  p.dkp = (int)DKP; p.kv_div = (int)kv_batch_div; p.scale_log2e = scale * 1.4426950408889634f; p.out_alpha = out_alpha; p.alpha2 = alpha2;
  const int nqt = (int)((Lq + BQ - 1) / BQ);
  // CTAs per (image, head): enough to fill the machine once (one CTA per SM), at most one per pair of query tiles.  More, smaller CTAs were
  // measured slower (round 2, call K: 4 CTAs per pair for a fuller last wave: 94 vs 83 us at 32 x 8 x 4096) - the per-CTA prologue (TMEM
  // allocation, barrier init, context load) is not small against ~30 us of work per CTA
  int64_t gx = ((int64_t)fyc_sm_count() + heads * NB - 1) / (heads * NB);
  if (gx < 1) gx = 1;
  if (gx > (nqt + 1) / 2) gx = (nqt + 1) / 2;
  dim3 grid((unsigned)gx, (unsigned)heads, (unsigned)NB);
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 40) {
    using Cfg = CxCfg<1, 48>;
    auto kern = attention_cx_kernel<1, 48, 0>;
    static bool attr = false;
    if (!attr) { FYC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr = true; }
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(mq, mk, mv, mk2, mv2, p);
  } else {
    using Cfg = CxCfg<2, 80>;
    auto kern = attention_cx_kernel<2, 80, 0>;
    static bool attr = false;
    if (!attr) { FYC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)); attr = true; }
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(mq, mk, mv, mk2, mv2, p);
  }
  FYC_LAUNCH_CHECK();
  return FYC_OK;
}
