// GroupNorm (cross-frame or per-frame statistics) and LayerNorm.  Both are HBM-bound: one read for the
// statistics, one read + one write for the apply; statistics are accumulated in fp32 per thread, combined in
// fp64 across CTAs so that E[x^2]-E[x]^2 does not cancel.
#include <stdlib.h>

#include "common.cuh"

// FYC_ZIGZAG (default 1): the norm kernels walk their input back to front.  Activations at the 64x64 level are 84 MB, the L2 is 126
// MB: a consumer that starts where its producer stopped finds the most recently written half still cached, one that starts at the
// front finds nothing.  GEMM / conv / attention write front to back, so the norms between them run back to front.
static int fyc_zigzag() {
  const char* e = getenv("FYC_ZIGZAG");
  return (e && e[0] == '0') ? 0 : 1;
}

// Waves of resident statistics CTAs (4 per SM) the GroupNorm statistics pass is cut into (A/B switch FYC_GN_WAVES=1|2|3).  Round 1 used 3
// (841 CTAs of 78 rows per clip at level 0: 13 rows per thread - the stream ends before the 8-deep load pipeline pays, and every CTA has
// its shared-memory reduction and partial write); same-box A/B (round 2, call N): GroupNorm per UNet forward 3.37 / 3.14 / 3.17 ms and per
// VAE decode 8.89 / 8.69 / 8.89 ms for 3 / 2 / 1 waves.
static int fyc_gn_waves() {
  static int w = -1;
  if (w < 0) {
    const char* e = getenv("FYC_GN_WAVES");
    w = (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 2;
  }
  return w;
}

namespace {
// Blackwell packed fp32 pairs (FFMA2 / FADD2 / FMUL2): one issue slot for two lanes' worth of work.  The norm kernels were
// instruction-bound (ncu: 18-20 issued instructions per element at ~50 % issue utilisation, 2.5-3 TB/s), not memory-bound.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 p, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// one 32-bit word holding two bf16 -> packed fp32 pair (element 0 in the low half)
__device__ __forceinline__ f32x2 bf2_to_f2(uint32_t w) { return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ uint32_t f2_to_bf2(f32x2 p) {
  float lo, hi; upk2(p, lo, hi);
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
}  // namespace

// ---------------------------------------------------------------------------------------------------------
// stats: x viewed as [NB, R, C]; grid (chunks, NB); each CTA reduces rows [r0, r1) for all channels and writes ONE
// partial (sum, sumsq) per group.  No atomics anywhere: the per-thread channel sums are combined through shared
// memory in a fixed order and the per-CTA partials are summed in chunk order by gn_finalize_kernel, so the
// statistics (and therefore the whole engine) are bit-reproducible run to run.
// Two-source form (x2 != nullptr): the normalised tensor is the channel concatenation [x (C1 channels) | x2 (C - C1 channels)] of two
// tensors that are never concatenated in memory - the skip connections of the up blocks (torch.cat at unet_blocks.py:763,885 followed
// by ResnetBlock3D.norm1).  A thread's channel vector lies wholly in one source (C1 % V == 0).
template <typename T, int V>
__global__ void __launch_bounds__(256, 4) gn_stats_kernel(const T* __restrict__ x, float2* __restrict__ partials, int64_t R,
                                                       int C, int G, int64_t rows_per_cta, int rev, const T* __restrict__ x2, int C1) {
  extern __shared__ float s_ch[];   // [RY][C][2]
  const int cpg = C / G;
  const int cvn = C / V;
  const int TX = cvn < 256 ? cvn : 256;
  const int RY = 256 / TX;
  const int tx = threadIdx.x % TX, ry = threadIdx.x / TX;
  // rev: walk the tensor back to front (zig-zag against the producer, which wrote it front to back: its tail is what L2 still holds)
  const int64_t nb = rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int64_t bx = rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int64_t r0 = bx * rows_per_cta;
  const int64_t r1 = (r0 + rows_per_cta < R) ? r0 + rows_per_cta : R;
  if (ry < RY) {
    for (int cv = tx; cv < cvn; cv += TX) {
      // row pointer and row stride of this channel vector's source; below `base + r * C + cv * V` addresses row r
      const bool second = x2 != nullptr && cv * V >= C1;
      const int ldx = x2 == nullptr ? C : (second ? C - C1 : C1);
      const T* base0 = second ? x2 + nb * R * ldx + (cv * V - C1) : x + nb * R * ldx + cv * V;
      float s[V], q[V];
#pragma unroll
      for (int e = 0; e < V; ++e) { s[e] = 0.f; q[e] = 0.f; }
      int64_t r = r0 + ry;
      if constexpr (sizeof(T) == 2 && V == 8) {
        // bf16: 8 independent 16-byte loads in flight per thread, kept packed until they are summed (the 4-deep version was
        // latency-bound at 2.6 TB/s: long_sb stalls with 45 % of the warps resident, profiles/round1_misc_full.md)
        for (; r + 7 * RY < r1; r += 8 * RY) {
          uint4 raw[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) raw[u] = __ldg(reinterpret_cast<const uint4*>(base0 + (r + (int64_t)u * RY) * ldx));
          f32x2 sp[4], qp[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) { sp[e] = pk2(s[2 * e], s[2 * e + 1]); qp[e] = pk2(q[2 * e], q[2 * e + 1]); }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { const f32x2 v2 = bf2_to_f2(w[e]); sp[e] = add2(sp[e], v2); qp[e] = fma2(v2, v2, qp[e]); }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) { upk2(sp[e], s[2 * e], s[2 * e + 1]); upk2(qp[e], q[2 * e], q[2 * e + 1]); }
        }
      }
      for (; r + 3 * RY < r1; r += 4 * RY) {      // 4 independent 16-byte loads in flight per thread
        float f[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const T* src = base0 + (r + (int64_t)u * RY) * ldx;
          if constexpr (V == 8) Vec8<T>::load(src, f[u]);
          else if constexpr (V == 4) Vec4<T>::load(src, f[u]);
          else f[u][0] = to_f(*src);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int e = 0; e < V; ++e) { s[e] += f[u][e]; q[e] = fmaf(f[u][e], f[u][e], q[e]); }
      }
      for (; r < r1; r += RY) {
        float f[8];
        if constexpr (V == 8) Vec8<T>::load(base0 + r * ldx, f);
        else if constexpr (V == 4) Vec4<T>::load(base0 + r * ldx, f);
        else f[0] = to_f(base0[r * ldx]);
#pragma unroll
        for (int e = 0; e < V; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        s_ch[((size_t)ry * C + cv * V + e) * 2] = s[e];
        s_ch[((size_t)ry * C + cv * V + e) * 2 + 1] = q[e];
      }
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += 256) {
    float as = 0.f, aq = 0.f;
    for (int y = 0; y < RY; ++y)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) { as += s_ch[((size_t)y * C + c) * 2]; aq += s_ch[((size_t)y * C + c) * 2 + 1]; }
    partials[((int64_t)nb * gridDim.x + bx) * G + g] = make_float2(as, aq);      // slot = logical chunk: the finalize order is unchanged
  }
}

// finalize: one CTA per (group, nb).  mean/rstd of the group from the chunk partials (fp64, fixed reduction tree), then per channel
// scale = rstd * gamma, shift = beta - mean * scale   (same form as ATen's CPU kernel).  (One CTA per nb walked all chunks x groups
// with 8 warps: 26 us for the cross-frame case NB = 2 - as long as the statistics pass itself.)
__global__ void __launch_bounds__(128) gn_finalize_kernel(const float2* __restrict__ partials, int chunks,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ scale, float* __restrict__ shift, int C, int G,
                                                          double count, float eps) {
  __shared__ double s_red[2][4];
  __shared__ float s_stat[2];
  const int g = blockIdx.x;
  const int64_t nb = blockIdx.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double s = 0.0, q = 0.0;
  for (int k = threadIdx.x; k < chunks; k += 128) { const float2 p = partials[(nb * chunks + k) * G + g]; s += (double)p.x; q += (double)p.y; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if (lane == 0) { s_red[0][wid] = s; s_red[1][wid] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double st = (s_red[0][0] + s_red[0][1]) + (s_red[0][2] + s_red[0][3]);
    const double qt = (s_red[1][0] + s_red[1][1]) + (s_red[1][2] + s_red[1][3]);
    const double mean = st / count;
    double var = qt / count - mean * mean;
    if (var < 0) var = 0;
    s_stat[0] = (float)mean;
    s_stat[1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int cpg = C / G;
  for (int c = g * cpg + threadIdx.x; c < (g + 1) * cpg; c += 128) {
    const float sc = s_stat[1] * gamma[c];
    scale[nb * C + c] = sc;
    shift[nb * C + c] = beta[c] - s_stat[0] * sc;
  }
}

template <typename T, int V, bool SILU>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, T* __restrict__ out, int64_t R,
                                                       int C, int64_t total_vec, const T* __restrict__ x2, int C1) {
  const int cvn = C / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;     // vectors in flight per thread
  // (channel vector, row) of vector i0 and of one grid stride, so that the loop advances them by addition: the 64-bit
  // div / mod per vector of the first version cost more issue slots than the loads
  const int64_t i00 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int cv = (int)(i00 % cvn);
  int64_t row = i00 / cvn;            // row over all NB images
  const int dcv = (int)(stride % cvn);
  const int64_t drow = stride / cvn;
  for (int64_t i0 = i00; i0 < total_vec; i0 += stride * U) {
    float f[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total_vec) {
        const T* src = x + i * V;
        if (x2 != nullptr) {                      // two-source form: (row, channel vector) of vector i -> its source tensor
          const int64_t rw = i / cvn; const int c = (int)(i - rw * cvn) * V;
          src = c < C1 ? x + rw * C1 + c : x2 + rw * (C - C1) + (c - C1);
        }
        if constexpr (V == 8) Vec8<T>::load(src, f[u]);
        else if constexpr (V == 4) Vec4<T>::load(src, f[u]);
        else f[u][0] = to_f(*src);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total_vec) {
        const int64_t nb = (int64_t)((uint32_t)row / (uint32_t)R);      // NB * R < 2^31 (checked by the caller)
        const float* sc = scale + nb * C + cv * V;
        const float* sh = shift + nb * C + cv * V;
        float scv[V], shv[V];
        if constexpr (V >= 4) {     // 16-byte parameter loads (scalar loads saturated the LSU queue: lg_throttle)
#pragma unroll
          for (int e = 0; e < V; e += 4) {
            float4 a = __ldg(reinterpret_cast<const float4*>(sc + e));
            float4 b = __ldg(reinterpret_cast<const float4*>(sh + e));
            scv[e] = a.x; scv[e + 1] = a.y; scv[e + 2] = a.z; scv[e + 3] = a.w;
            shv[e] = b.x; shv[e + 1] = b.y; shv[e + 2] = b.z; shv[e + 3] = b.w;
          }
        } else {
          scv[0] = __ldg(sc); shv[0] = __ldg(sh);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
          float y = fmaf(f[u][e], scv[e], shv[e]);
          f[u][e] = SILU ? (sizeof(T) == 2 ? silu_fast(y) : silu_f(y)) : y;
        }
        if constexpr (V == 8) Vec8<T>::store(out + i * V, f[u]);
        else if constexpr (V == 4) Vec4<T>::store(out + i * V, f[u]);
        else out[i] = from_f<T>(f[u][0]);
      }
      cv += dcv; row += drow;
      if (cv >= cvn) { cv -= cvn; ++row; }
    }
  }
}

// bf16 apply, second form: a thread owns ONE 8-channel vector (scale / shift live in 16 registers, loaded once) and walks rows,
// U rows in flight; the math is packed fp32x2.  The grid-stride form above re-derived (image, channel) and re-loaded four
// parameter vectors for every 16 bytes of data: 160 issued instructions per vector, 51 us for an 84 MB tensor.
// [r2] Row blocks (RY x U consecutive rows) are dealt to the CTAs round-robin - neighbouring CTAs stream neighbouring memory at the same
// time instead of 1184 far-apart private chunks - and the NEXT block's loads are issued before the current block's SiLU math (register
// double buffer), so a CTA always has 2 x U x 16 bytes per thread in flight.  SiLU = y (0.5 + 0.5 tanh(y / 2)): one MUFU.TANH per
// element instead of EX2 + RCP (the apply kernel spent a third of its issue slots and all of its MUFU slots there).
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <bool SILU>
__global__ void __launch_bounds__(256, 3) gn_apply_rows_kernel(const bf16* __restrict__ x, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, bf16* __restrict__ out, int64_t R,
                                                            int C, int64_t rows_per_cta, const bf16* __restrict__ x2, int C1) {
  constexpr int U = 4;
  const int cvn = C / 8;
  const int TX = cvn < 256 ? cvn : 256;
  const int RY = 256 / TX;
  const int tx = threadIdx.x % TX, ry = threadIdx.x / TX;
  if (ry >= RY) return;
  const int64_t nb = blockIdx.y;
  const int64_t blk_rows = (int64_t)RY * U;                     // rows one CTA iteration covers
  const int64_t nblk = (R + blk_rows - 1) / blk_rows;
  (void)rows_per_cta;
  bf16* ob = out + nb * R * C;
  for (int cv = tx; cv < cvn; cv += TX) {
    const bool second = x2 != nullptr && cv * 8 >= C1;
    const int ldx = x2 == nullptr ? C : (second ? C - C1 : C1);
    const bf16* xb = second ? x2 + nb * R * ldx + (cv * 8 - C1) : x + nb * R * ldx + cv * 8;     // row r of this channel vector: xb + r * ldx
    f32x2 sc[4], sh[4];
    {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale + nb * C + cv * 8)), a1 = __ldg(reinterpret_cast<const float4*>(scale + nb * C + cv * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + nb * C + cv * 8)), b1 = __ldg(reinterpret_cast<const float4*>(shift + nb * C + cv * 8 + 4));
      sc[0] = pk2(a0.x, a0.y); sc[1] = pk2(a0.z, a0.w); sc[2] = pk2(a1.x, a1.y); sc[3] = pk2(a1.z, a1.w);
      sh[0] = pk2(b0.x, b0.y); sh[1] = pk2(b0.z, b0.w); sh[2] = pk2(b1.x, b1.y); sh[3] = pk2(b1.z, b1.w);
    }
    const f32x2 half2 = pk2(0.5f, 0.5f);
    auto load_blk = [&](int64_t blk, uint4* raw) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = blk * blk_rows + (int64_t)u * RY + ry;
        raw[u] = make_uint4(0, 0, 0, 0);
        if (blk < nblk && rr < R) raw[u] = __ldg(reinterpret_cast<const uint4*>(xb + rr * ldx));
      }
    };
    uint4 cur[U], nxt[U];
    int64_t blk = blockIdx.x;
    load_blk(blk, cur);
    for (; blk < nblk; blk += gridDim.x) {
      load_blk(blk + gridDim.x, nxt);                             // next block's rows are in flight during this block's math
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = blk * blk_rows + (int64_t)u * RY + ry;
        if (rr >= R) break;
        const uint32_t w[4] = {cur[u].x, cur[u].y, cur[u].z, cur[u].w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          f32x2 y = fma2(bf2_to_f2(w[e]), sc[e], sh[e]);
          if (SILU) {                       // y * sigmoid(y) = y * (0.5 + 0.5 tanh(y / 2))
            float h0, h1; upk2(mul2(y, half2), h0, h1);
            y = mul2(y, fma2(pk2(tanh_fast(h0), tanh_fast(h1)), half2, half2));
          }
          o[e] = f2_to_bf2(y);
        }
        *reinterpret_cast<uint4*>(ob + rr * C + cv * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) cur[u] = nxt[u];
    }
  }
}

static int64_t gn_max_chunks(int64_t NB) { return ((int64_t)fyc_sm_count() * 12 + NB - 1) / NB + 1; }   // ~3 waves of 4 CTAs per SM
static int64_t gn_partials(int64_t NB, int64_t G) { return (NB * gn_max_chunks(NB) * G + 1) / 2 * 2; }   // even: keeps scale/shift 16-byte aligned

extern "C" size_t fyc_groupnorm_workspace_bytes(int64_t NB, int64_t C, int64_t G) {
  return (size_t)(gn_partials(NB, G) * sizeof(float2) + NB * C * 2 * sizeof(float));
}

template <typename T, int V>
static int32_t groupnorm_impl(const T* x, const float* gamma, const float* beta, T* out, int64_t NB, int64_t R, int C,
                              int G, float eps, int silu, void* ws, cudaStream_t st, const T* x2 = nullptr, int C1 = 0) {
  float2* partials = (float2*)ws;
  float* scale = (float*)(partials + gn_partials(NB, G));
  float* shift = scale + NB * C;
  const int cvn = C / V;
  const int TX = cvn < 256 ? cvn : 256;
  const int RY = 256 / TX;
  int64_t target = ceil_div64((int64_t)fyc_sm_count() * 4 * fyc_gn_waves(), NB);   // CTAs per nb (<= gn_max_chunks(NB) - 1, the workspace bound)
  int64_t rows_per_cta = ceil_div64(R, target);
  if (rows_per_cta < 8 * RY) rows_per_cta = 8 * RY;
  rows_per_cta = ceil_div64(rows_per_cta, RY) * RY;
  const int chunks = (int)ceil_div64(R, rows_per_cta);
  FYC_CHECK(chunks <= gn_max_chunks(NB), "groupnorm: internal chunk count");
  const size_t smem = (size_t)RY * C * 2 * sizeof(float);
  FYC_CHECK(smem <= 200 * 1024, "groupnorm: C=%d too large", C);
  auto kern = gn_stats_kernel<T, V>;
  if (smem > 48 * 1024) FYC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)chunks, (unsigned)NB);
  kern<<<grid, 256, smem, st>>>(x, partials, R, C, G, rows_per_cta, fyc_zigzag(), x2, C1);
  FYC_LAUNCH_CHECK();
  gn_finalize_kernel<<<dim3((unsigned)G, (unsigned)NB), 128, 0, st>>>(partials, chunks, gamma, beta, scale, shift, C, G,
                                                                       (double)R * (C / G), eps);
  FYC_LAUNCH_CHECK();
  if constexpr (sizeof(T) == 2 && V == 8) {
    // one wave of resident CTAs (3 per SM); each walks the row blocks (RY x 4 rows) round-robin
    int64_t want = ceil_div64((int64_t)fyc_sm_count() * 3, NB);
    const int64_t nblk = ceil_div64(R, (int64_t)RY * 4);
    if (want > nblk) want = nblk;
    const int64_t rpc = 0;
    dim3 ga((unsigned)want, (unsigned)NB);
    if (silu) gn_apply_rows_kernel<true><<<ga, 256, 0, st>>>((const bf16*)x, scale, shift, (bf16*)out, R, C, rpc, (const bf16*)x2, C1);
    else gn_apply_rows_kernel<false><<<ga, 256, 0, st>>>((const bf16*)x, scale, shift, (bf16*)out, R, C, rpc, (const bf16*)x2, C1);
    FYC_LAUNCH_CHECK();
    return FYC_OK;
  }
  int64_t total_vec = NB * R * cvn;
  int64_t blocks = ceil_div64(total_vec, 256);
  int64_t cap = (int64_t)fyc_sm_count() * 16;
  unsigned gb = (unsigned)(blocks > cap ? cap : blocks);
  if (silu) gn_apply_kernel<T, V, true><<<gb, 256, 0, st>>>(x, scale, shift, out, R, C, total_vec, x2, C1);
  else gn_apply_kernel<T, V, false><<<gb, 256, 0, st>>>(x, scale, shift, out, R, C, total_vec, x2, C1);
  FYC_LAUNCH_CHECK();
  return FYC_OK;
}

extern "C" int32_t fyc_groupnorm(const void* x, const float* gamma, const float* beta, void* out, int64_t NB, int64_t R,
                                 int64_t C, int64_t G, float eps, int32_t silu, int32_t dtype, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  FYC_CHECK(G > 0 && C % G == 0, "groupnorm: C=%lld not divisible by G=%lld", (long long)C, (long long)G);
  FYC_CHECK(workspace && workspace_bytes >= fyc_groupnorm_workspace_bytes(NB, C, G), "groupnorm: workspace too small");
  FYC_CHECK(NB > 0 && NB < 65536 && R > 0 && C < (1 << 20) && NB * R < (1ll << 31), "groupnorm: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FYC_BF16) {
    if (C % 8 == 0) return groupnorm_impl<bf16, 8>((const bf16*)x, gamma, beta, (bf16*)out, NB, R, (int)C, (int)G, eps, silu, workspace, st);
    return groupnorm_impl<bf16, 1>((const bf16*)x, gamma, beta, (bf16*)out, NB, R, (int)C, (int)G, eps, silu, workspace, st);
  } else if (dtype == FYC_F32) {
    if (C % 4 == 0) return groupnorm_impl<float, 4>((const float*)x, gamma, beta, (float*)out, NB, R, (int)C, (int)G, eps, silu, workspace, st);
    return groupnorm_impl<float, 1>((const float*)x, gamma, beta, (float*)out, NB, R, (int)C, (int)G, eps, silu, workspace, st);
  }
  FYC_CHECK(false, "groupnorm: unknown dtype %d", dtype);
}

extern "C" int32_t fyc_groupnorm_concat(const void* x1, int64_t C1, const void* x2, int64_t C2, const float* gamma, const float* beta,
                                        void* out, int64_t NB, int64_t R, int64_t G, float eps, int32_t silu, int32_t dtype,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  const int64_t C = C1 + C2;
  FYC_CHECK(x1 && x2 && C1 > 0 && C2 > 0, "groupnorm_concat: bad arguments");
  FYC_CHECK(G > 0 && C % G == 0, "groupnorm_concat: C=%lld not divisible by G=%lld", (long long)C, (long long)G);
  FYC_CHECK(workspace && workspace_bytes >= fyc_groupnorm_workspace_bytes(NB, C, G), "groupnorm_concat: workspace too small");
  FYC_CHECK(NB > 0 && NB < 65536 && R > 0 && C < (1 << 20) && NB * R < (1ll << 31), "groupnorm_concat: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FYC_BF16) {
    FYC_CHECK(C1 % 8 == 0 && C2 % 8 == 0 && ((((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)out)) & 15) == 0, "groupnorm_concat(bf16): channel counts must be multiples of 8, pointers 16-byte aligned");
    return groupnorm_impl<bf16, 8>((const bf16*)x1, gamma, beta, (bf16*)out, NB, R, (int)C, (int)G, eps, silu, workspace, st, (const bf16*)x2, (int)C1);
  } else if (dtype == FYC_F32) {
    FYC_CHECK(C1 % 4 == 0 && C2 % 4 == 0, "groupnorm_concat(f32): channel counts must be multiples of 4");
    return groupnorm_impl<float, 4>((const float*)x1, gamma, beta, (float*)out, NB, R, (int)C, (int)G, eps, silu, workspace, st, (const float*)x2, (int)C1);
  }
  FYC_CHECK(false, "groupnorm_concat: unknown dtype %d", dtype);
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers (two-pass mean / centred variance), optional PE add.
template <typename T, int V, int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, T* __restrict__ out, int64_t M,
                                                        int C, float eps, const float* __restrict__ pe,
                                                        int64_t rows_per_frame, int64_t frames) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int cvn = C / V;
  const T* xr = x + row * C;
  float v[NV][V];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int cv = lane + 32 * i;
    if (cv < cvn) {
      if constexpr (V == 8) Vec8<T>::load(xr + cv * V, v[i]); else Vec4<T>::load(xr + cv * V, v[i]);
#pragma unroll
      for (int e = 0; e < V; ++e) sum += v[i][e];
    }
  }
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int cv = lane + 32 * i;
    if (cv < cvn) {
#pragma unroll
      for (int e = 0; e < V; ++e) { float d = v[i][e] - mean; sq = fmaf(d, d, sq); }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
  const float* per = pe ? pe + ((row / rows_per_frame) % frames) * C : nullptr;
  T* orow = out + row * C;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int cv = lane + 32 * i;
    if (cv < cvn) {
      float o[V], gm[V], bt[V];
      // 16-byte parameter loads (scalar loads here saturated the LSU queue: lg_throttle in profiles/round1)
#pragma unroll
      for (int e = 0; e < V; e += 4) {
        float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + cv * V + e));
        float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + cv * V + e));
        gm[e] = g4.x; gm[e + 1] = g4.y; gm[e + 2] = g4.z; gm[e + 3] = g4.w;
        bt[e] = b4.x; bt[e + 1] = b4.y; bt[e + 2] = b4.z; bt[e + 3] = b4.w;
      }
      if (per) {
#pragma unroll
        for (int e = 0; e < V; e += 4) {
          float4 p4 = __ldg(reinterpret_cast<const float4*>(per + cv * V + e));
          bt[e] += p4.x; bt[e + 1] += p4.y; bt[e + 2] += p4.z; bt[e + 3] += p4.w;
        }
      }
#pragma unroll
      for (int e = 0; e < V; ++e) o[e] = (v[i][e] - mean) * rstd * gm[e] + bt[e];
      if constexpr (V == 8) Vec8<T>::store(orow + cv * V, o); else Vec4<T>::store(orow + cv * V, o);
    }
  }
}

// bf16 LayerNorm, RPW rows per warp: every row's 16-byte vectors are requested before any is used (RPW x NV loads in
// flight per lane instead of NV), gamma / beta are read once per warp instead of once per row.  The one-row kernel ran at
// 2.6 TB/s of the 6.6 TB/s the copy benchmark reaches: too few bytes in flight per SM and ~5 parameter loads per data load.
template <int NV, int RPW>
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, bf16* __restrict__ out, int64_t M,
                                                             int C, float eps, const float* __restrict__ pe,
                                                             int64_t rows_per_frame, int64_t frames) {
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  if (row0 >= M) return;
  const int cvn = C / 8;
  uint4 raw[RPW][NV];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int cv = lane + 32 * i;
      raw[r][i] = make_uint4(0, 0, 0, 0);
      if (cv < cvn && row0 + r < M) raw[r][i] = __ldg(reinterpret_cast<const uint4*>(x + (row0 + r) * C + cv * 8));
    }
  float gm[NV][8], bt[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int cv = lane + 32 * i;
    if (cv < cvn) {
#pragma unroll
      for (int e = 0; e < 8; e += 4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8 + e));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8 + e));
        gm[i][e] = g4.x; gm[i][e + 1] = g4.y; gm[i][e + 2] = g4.z; gm[i][e + 3] = g4.w;
        bt[i][e] = b4.x; bt[i][e + 1] = b4.y; bt[i][e + 2] = b4.z; bt[i][e + 3] = b4.w;
      }
    }
  }
  const float inv_c = 1.0f / (float)C;
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int64_t row = row0 + r;
    if (row >= M) break;                      // warp-uniform
    float v[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { v[i][2 * e] = __uint_as_float(w[e] << 16); v[i][2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
      if (lane + 32 * i < cvn) {
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += v[i][e];
      }
    }
    const float mean = warp_sum(sum) * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < cvn) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mean; sq = fmaf(d, d, sq); }
      }
    const float rstd = rsqrtf(warp_sum(sq) * inv_c + eps);
    const float* per = pe ? pe + ((row / rows_per_frame) % frames) * C : nullptr;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int cv = lane + 32 * i;
      if (cv < cvn) {
        float o[8], pb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (per) {
          const float4 p0 = __ldg(reinterpret_cast<const float4*>(per + cv * 8)), p1 = __ldg(reinterpret_cast<const float4*>(per + cv * 8 + 4));
          pb[0] = p0.x; pb[1] = p0.y; pb[2] = p0.z; pb[3] = p0.w; pb[4] = p1.x; pb[5] = p1.y; pb[6] = p1.z; pb[7] = p1.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * gm[i][e] + (bt[i][e] + pb[e]);
        Vec8<bf16>::store(out + row * C + cv * 8, o);
      }
    }
  }
}

// bf16 LayerNorm for C = 40 * LPR (320 / 640 / 1280): LPR lanes share a row, five 16-byte vectors per lane - every lane is busy
// (the warp-per-row kernels idle 24 of 32 lanes on the second vector of a 320-wide row), 32 / LPR rows per warp pass, PASSES passes
// with all loads of a pass issued up front.  gamma / beta stay in registers for the warp's lifetime; the arithmetic is the same
// two-pass (mean, then centred variance) as the reference kernel, on packed fp32 pairs: ~5 issued instructions per element
// instead of ~18 (ncu, profiles/round1_norms.md).
template <int LPR, int PASSES, bool HAS_PE>
__global__ void __launch_bounds__(256, 2) layernorm_lpr_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, bf16* __restrict__ out, int64_t M,
                                                               float eps, const float* __restrict__ pe, int64_t rows_per_frame,
                                                               int64_t frames, int rev) {
  constexpr int C = LPR * 40, RPP = 32 / LPR;           // channels; rows per warp pass
  constexpr int VS = LPR * 8;                           // element stride between a lane's consecutive vectors
  const int lane = threadIdx.x & 31, sub = lane % LPR, rr = lane / LPR;
  // rev: CTAs take the row blocks back to front - the GEMM that produced x wrote it front to back (its tail is still in L2) and the
  // GEMM that consumes `out` reads front to back (the front is what this kernel then wrote last)
  const int64_t bxl = rev ? (int64_t)gridDim.x - 1 - blockIdx.x : (int64_t)blockIdx.x;
  const int64_t row_base = (bxl * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (RPP * PASSES);
  if (row_base >= M) return;
  f32x2 gm[5][4];                                       // beta is re-read from L1 per vector: 40 more registers would halve the occupancy
  const float* gp = gamma + sub * 8;
  const float* bp = beta + sub * 8;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp + i * VS)), g1 = __ldg(reinterpret_cast<const float4*>(gp + i * VS + 4));
    gm[i][0] = pk2(g0.x, g0.y); gm[i][1] = pk2(g0.z, g0.w); gm[i][2] = pk2(g1.x, g1.y); gm[i][3] = pk2(g1.z, g1.w);
  }
  const float inv_c = 1.0f / (float)C;
#pragma unroll 1
  for (int ps = 0; ps < PASSES; ++ps) {
    const int64_t row = row_base + ps * RPP + rr;
    const bool ok = row < M;
    const int64_t rowc = ok ? row : M - 1;            // out-of-range lanes shadow the last row (shuffles stay full-warp), never store
    const bf16* xr = x + rowc * C + sub * 8;
    uint4 raw[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) raw[i] = __ldg(reinterpret_cast<const uint4*>(xr + i * VS));
    f32x2 v[5][4];
    f32x2 s2 = pk2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { v[i][e] = bf2_to_f2(w[e]); s2 = add2(s2, v[i][e]); }
    }
    float s0, s1; upk2(s2, s0, s1);
    float sum = s0 + s1;
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    const f32x2 nmean = pk2(-mean, -mean);
    f32x2 q2 = pk2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) { v[i][e] = add2(v[i][e], nmean); q2 = fma2(v[i][e], v[i][e], q2); }
    float q0, q1; upk2(q2, q0, q1);
    float sq = q0 + q1;
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
    const f32x2 rstd2 = pk2(rstd, rstd);
    const float* pp = HAS_PE ? pe + ((rowc / rows_per_frame) % frames) * C + sub * 8 : nullptr;
    bf16* orow = out + rowc * C + sub * 8;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp + i * VS)), b1 = __ldg(reinterpret_cast<const float4*>(bp + i * VS + 4));
      f32x2 bb[4] = {pk2(b0.x, b0.y), pk2(b0.z, b0.w), pk2(b1.x, b1.y), pk2(b1.z, b1.w)};
      if (HAS_PE) {
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(pp + i * VS)), p1 = __ldg(reinterpret_cast<const float4*>(pp + i * VS + 4));
        bb[0] = add2(bb[0], pk2(p0.x, p0.y)); bb[1] = add2(bb[1], pk2(p0.z, p0.w));
        bb[2] = add2(bb[2], pk2(p1.x, p1.y)); bb[3] = add2(bb[3], pk2(p1.z, p1.w));
      }
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = f2_to_bf2(fma2(mul2(v[i][e], rstd2), gm[i][e], bb[e]));
      if (ok) *reinterpret_cast<uint4*>(orow + i * VS) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// LayerNorm STATISTICS only (fyc_layernorm_stats): per row rstd (fp32) and the 8-column bf16 "aug" row [m_hi, m_hi, m_lo, m_lo, 0, 0, 0, 0]
// (mean = m_hi + m_lo) that the LN-folded GEMM appends to its K dimension (fyc.h FYC_EPI_LNFOLD).  Same lane layout and the same
// two-pass arithmetic as layernorm_lpr_kernel - one read of x, no write of a normalised copy.  PASSES x 5 independent 16-byte loads
// per lane are requested before the first is used.
__device__ __forceinline__ void ln_write_stats(float* __restrict__ rstd_out, bf16* __restrict__ aug, int64_t row, float mean, float rstd) {
  rstd_out[row] = rstd;
  if (aug == nullptr) return;
  const bf16 hi = __float2bfloat16_rn(mean);
  const bf16 lo = __float2bfloat16_rn(mean - __bfloat162float(hi));
  const uint32_t hh = (uint32_t)__bfloat16_as_ushort(hi) * 0x10001u, ll = (uint32_t)__bfloat16_as_ushort(lo) * 0x10001u;
  *reinterpret_cast<uint4*>(aug + row * 8) = make_uint4(hh, ll, 0u, 0u);
}

template <int LPR, int PASSES>
__global__ void __launch_bounds__(256, 2) ln_stats_lpr_kernel(const bf16* __restrict__ x, float* __restrict__ rstd_out, bf16* __restrict__ aug,
                                                              int64_t M, float eps, int rev) {
  constexpr int C = LPR * 40, RPP = 32 / LPR, VS = LPR * 8;
  const int lane = threadIdx.x & 31, sub = lane % LPR, rr = lane / LPR;
  const int64_t bxl = rev ? (int64_t)gridDim.x - 1 - blockIdx.x : (int64_t)blockIdx.x;
  const int64_t row_base = (bxl * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (RPP * PASSES);
  if (row_base >= M) return;
  const float inv_c = 1.0f / (float)C;
  uint4 raw[PASSES][5];
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    const int64_t row = row_base + ps * RPP + rr;
    const int64_t rowc = row < M ? row : M - 1;
    const bf16* xr = x + rowc * C + sub * 8;
#pragma unroll
    for (int i = 0; i < 5; ++i) raw[ps][i] = __ldg(reinterpret_cast<const uint4*>(xr + i * VS));
  }
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    const int64_t row = row_base + ps * RPP + rr;
    f32x2 v[5][4];
    f32x2 s2 = pk2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const uint32_t w[4] = {raw[ps][i].x, raw[ps][i].y, raw[ps][i].z, raw[ps][i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { v[i][e] = bf2_to_f2(w[e]); s2 = add2(s2, v[i][e]); }
    }
    float s0, s1; upk2(s2, s0, s1);
    float sum = s0 + s1;
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    const f32x2 nmean = pk2(-mean, -mean);
    f32x2 q2 = pk2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) { const f32x2 d = add2(v[i][e], nmean); q2 = fma2(d, d, q2); }
    float q0, q1; upk2(q2, q0, q1);
    float sq = q0 + q1;
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
    if (sub == 0 && row < M) ln_write_stats(rstd_out, aug, row, mean, rstd);
  }
}

// generic widths (C % 8 == 0, C <= 2048; bf16) and fp32 rows: one warp per row
template <typename T, int V, int NV>
__global__ void __launch_bounds__(256) ln_stats_kernel(const T* __restrict__ x, float* __restrict__ rstd_out, bf16* __restrict__ aug, int64_t M, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int cvn = C / V;
  const T* xr = x + row * C;
  float v[NV][V];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int cv = lane + 32 * i;
    if (cv < cvn) {
      if constexpr (V == 8) Vec8<T>::load(xr + cv * V, v[i]); else Vec4<T>::load(xr + cv * V, v[i]);
#pragma unroll
      for (int e = 0; e < V; ++e) sum += v[i][e];
    }
  }
  const float mean = warp_sum(sum) / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (lane + 32 * i < cvn) {
#pragma unroll
      for (int e = 0; e < V; ++e) { const float d = v[i][e] - mean; sq = fmaf(d, d, sq); }
    }
  const float rstd = rsqrtf(warp_sum(sq) / (float)C + eps);
  if (lane == 0) ln_write_stats(rstd_out, aug, row, mean, rstd);
}

extern "C" int32_t fyc_layernorm_stats(const void* x, float* rstd, void* aug, int64_t M, int64_t C, float eps, int32_t dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  FYC_CHECK(x && rstd && M > 0 && C > 0, "layernorm_stats: bad arguments");
  FYC_CHECK((((uintptr_t)x | (uintptr_t)aug) & 15) == 0 && (((uintptr_t)rstd) & 3) == 0, "layernorm_stats: alignment");
  bf16* ao = (bf16*)aug;
  const unsigned grid = (unsigned)ceil_div64(M, 8);
  if (dtype == FYC_BF16) {
    FYC_CHECK(C % 8 == 0 && C <= 2048, "layernorm_stats(bf16): C=%lld must be a multiple of 8 and <= 2048", (long long)C);
    const bf16* xb = (const bf16*)x;
    constexpr int PASSES = 4;
    if (C == 320) ln_stats_lpr_kernel<8, PASSES><<<(unsigned)ceil_div64(M, 8 * 4 * PASSES), 256, 0, st>>>(xb, rstd, ao, M, eps, fyc_zigzag());
    else if (C == 640) ln_stats_lpr_kernel<16, PASSES><<<(unsigned)ceil_div64(M, 8 * 2 * PASSES), 256, 0, st>>>(xb, rstd, ao, M, eps, fyc_zigzag());
    else if (C == 1280) ln_stats_lpr_kernel<32, PASSES><<<(unsigned)ceil_div64(M, 8 * 1 * PASSES), 256, 0, st>>>(xb, rstd, ao, M, eps, fyc_zigzag());
    else if (C <= 8 * 32 * 5) ln_stats_kernel<bf16, 8, 5><<<grid, 256, 0, st>>>(xb, rstd, ao, M, (int)C, eps);
    else ln_stats_kernel<bf16, 8, 8><<<grid, 256, 0, st>>>(xb, rstd, ao, M, (int)C, eps);
  } else if (dtype == FYC_F32) {
    FYC_CHECK(C % 4 == 0 && C <= 2048, "layernorm_stats(f32): C=%lld must be a multiple of 4 and <= 2048", (long long)C);
    if (C <= 4 * 32 * 5) ln_stats_kernel<float, 4, 5><<<grid, 256, 0, st>>>((const float*)x, rstd, ao, M, (int)C, eps);
    else if (C <= 4 * 32 * 10) ln_stats_kernel<float, 4, 10><<<grid, 256, 0, st>>>((const float*)x, rstd, ao, M, (int)C, eps);
    else ln_stats_kernel<float, 4, 16><<<grid, 256, 0, st>>>((const float*)x, rstd, ao, M, (int)C, eps);
  } else {
    FYC_CHECK(false, "layernorm_stats: unknown dtype %d", dtype);
  }
  FYC_LAUNCH_CHECK();
  return FYC_OK;
}

template <int LPR>
static void launch_ln_lpr(const bf16* xb, const float* gamma, const float* beta, bf16* ob, int64_t M, float eps, const float* pe,
                          int64_t rows_per_frame, int64_t frames, cudaStream_t st) {
  constexpr int PASSES = 2;
  const unsigned grid = (unsigned)ceil_div64(M, 8 * (32 / LPR) * PASSES);
  if (pe) layernorm_lpr_kernel<LPR, PASSES, true><<<grid, 256, 0, st>>>(xb, gamma, beta, ob, M, eps, pe, rows_per_frame, frames, fyc_zigzag());
  else layernorm_lpr_kernel<LPR, PASSES, false><<<grid, 256, 0, st>>>(xb, gamma, beta, ob, M, eps, pe, rows_per_frame, frames, fyc_zigzag());
}

extern "C" int32_t fyc_layernorm(const void* x, const float* gamma, const float* beta, void* out, int64_t M, int64_t C,
                                 float eps, const float* pe, int64_t rows_per_frame, int64_t frames, int32_t dtype,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  FYC_CHECK(M > 0 && C > 0, "layernorm: bad shape");
  if (pe) FYC_CHECK(rows_per_frame > 0 && frames > 0, "layernorm: pe needs rows_per_frame/frames");
  unsigned grid = (unsigned)ceil_div64(M, 8);
  if (dtype == FYC_BF16) {
    FYC_CHECK(C % 8 == 0 && C <= 8 * 32 * 8, "layernorm(bf16): C=%lld must be a multiple of 8 and <= 2048", (long long)C);
    FYC_CHECK((((uintptr_t)x | (uintptr_t)out) & 15) == 0 && (((uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)pe) & 15) == 0, "layernorm(bf16): 16-byte alignment");
    const bf16* xb = (const bf16*)x; bf16* ob = (bf16*)out;
    if (C == 320) launch_ln_lpr<8>(xb, gamma, beta, ob, M, eps, pe, rows_per_frame, frames, st);
    else if (C == 640) launch_ln_lpr<16>(xb, gamma, beta, ob, M, eps, pe, rows_per_frame, frames, st);
    else if (C == 1280) launch_ln_lpr<32>(xb, gamma, beta, ob, M, eps, pe, rows_per_frame, frames, st);
    else if (C <= 8 * 32 * 2) layernorm_bf16_kernel<2, 4><<<(unsigned)ceil_div64(M, 8 * 4), 256, 0, st>>>(xb, gamma, beta, ob, M, (int)C, eps, pe, rows_per_frame, frames);
    else if (C <= 8 * 32 * 3) layernorm_bf16_kernel<3, 2><<<(unsigned)ceil_div64(M, 8 * 2), 256, 0, st>>>(xb, gamma, beta, ob, M, (int)C, eps, pe, rows_per_frame, frames);
    else if (C <= 8 * 32 * 5) layernorm_kernel<bf16, 8, 5><<<grid, 256, 0, st>>>(xb, gamma, beta, ob, M, (int)C, eps, pe, rows_per_frame, frames);
    else layernorm_kernel<bf16, 8, 8><<<grid, 256, 0, st>>>(xb, gamma, beta, ob, M, (int)C, eps, pe, rows_per_frame, frames);
  } else if (dtype == FYC_F32) {
    FYC_CHECK(C % 4 == 0 && C <= 4 * 32 * 16, "layernorm(f32): C=%lld must be a multiple of 4 and <= 2048", (long long)C);
    if (C <= 4 * 32 * 5) layernorm_kernel<float, 4, 5><<<grid, 256, 0, st>>>((const float*)x, gamma, beta, (float*)out, M, (int)C, eps, pe, rows_per_frame, frames);
    else if (C <= 4 * 32 * 10) layernorm_kernel<float, 4, 10><<<grid, 256, 0, st>>>((const float*)x, gamma, beta, (float*)out, M, (int)C, eps, pe, rows_per_frame, frames);
    else layernorm_kernel<float, 4, 16><<<grid, 256, 0, st>>>((const float*)x, gamma, beta, (float*)out, M, (int)C, eps, pe, rows_per_frame, frames);
  } else {
    FYC_CHECK(false, "layernorm: unknown dtype %d", dtype);
  }
  FYC_LAUNCH_CHECK();
  return FYC_OK;
}
