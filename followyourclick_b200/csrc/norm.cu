// GroupNorm (cross-frame or per-frame statistics) and LayerNorm.  Both are HBM-bound: one read for the
// statistics, one read + one write for the apply; statistics are accumulated in fp32 per thread, combined in
// fp64 across CTAs so that E[x^2]-E[x]^2 does not cancel.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// stats: x viewed as [NB, R, C]; grid (chunks, NB); each CTA reduces rows [r0, r1) for all channels and writes ONE
// partial (sum, sumsq) per group.  No atomics anywhere: the per-thread channel sums are combined through shared
// memory in a fixed order and the per-CTA partials are summed in chunk order by gn_finalize_kernel, so the
// statistics (and therefore the whole engine) are bit-reproducible run to run.
template <typename T, int V>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, float2* __restrict__ partials, int64_t R,
                                                       int C, int G, int64_t rows_per_cta) {
  extern __shared__ float s_ch[];   // [RY][C][2]
  const int cpg = C / G;
  const int cvn = C / V;
  const int TX = cvn < 256 ? cvn : 256;
  const int RY = 256 / TX;
  const int tx = threadIdx.x % TX, ry = threadIdx.x / TX;
  const int64_t nb = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = (r0 + rows_per_cta < R) ? r0 + rows_per_cta : R;
  const T* base = x + nb * R * C;
  if (ry < RY) {
    for (int cv = tx; cv < cvn; cv += TX) {
      float s[V], q[V];
#pragma unroll
      for (int e = 0; e < V; ++e) { s[e] = 0.f; q[e] = 0.f; }
      int64_t r = r0 + ry;
      for (; r + 3 * RY < r1; r += 4 * RY) {      // 4 independent 16-byte loads in flight per thread
        float f[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const T* src = base + (r + (int64_t)u * RY) * C + cv * V;
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
        if constexpr (V == 8) Vec8<T>::load(base + r * C + cv * V, f);
        else if constexpr (V == 4) Vec4<T>::load(base + r * C + cv * V, f);
        else f[0] = to_f(base[r * C + cv]);
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
    partials[((int64_t)nb * gridDim.x + blockIdx.x) * G + g] = make_float2(as, aq);
  }
}

// finalize: one CTA per nb.  mean/rstd per group from the chunk partials (fp64, fixed order), then per channel
// scale = rstd * gamma, shift = beta - mean * scale   (same form as ATen's CPU kernel)
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float2* __restrict__ partials, int chunks,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ scale, float* __restrict__ shift, int C, int G,
                                                          double count, float eps) {
  extern __shared__ float s_stat[];   // [G][2] mean, rstd
  const int64_t nb = blockIdx.x;
  // one warp per group: lanes stride over the chunk partials, then a fixed-shape shuffle tree (deterministic)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int g = wid; g < G; g += nw) {
    double s = 0.0, q = 0.0;
    for (int k = lane; k < chunks; k += 32) { float2 p = partials[(nb * chunks + k) * G + g]; s += (double)p.x; q += (double)p.y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0) {
      double mean = s / count;
      double var = q / count - mean * mean;
      if (var < 0) var = 0;
      s_stat[2 * g] = (float)mean;
      s_stat[2 * g + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  __syncthreads();
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int g = c / cpg;
    float sc = s_stat[2 * g + 1] * gamma[c];
    scale[nb * C + c] = sc;
    shift[nb * C + c] = beta[c] - s_stat[2 * g] * sc;
  }
}

template <typename T, int V, bool SILU>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, T* __restrict__ out, int64_t R,
                                                       int C, int64_t total_vec) {
  const int cvn = C / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;     // vectors in flight per thread
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total_vec; i0 += stride * U) {
    float f[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total_vec) {
        if constexpr (V == 8) Vec8<T>::load(x + i * V, f[u]);
        else if constexpr (V == 4) Vec4<T>::load(x + i * V, f[u]);
        else f[u][0] = to_f(x[i]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= total_vec) break;
      const int cv = (int)(i % cvn);
      const int64_t nb = (i / cvn) / R;
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
  }
}

static int64_t gn_max_chunks(int64_t NB) { return ((int64_t)fyc_sm_count() * 4 + NB - 1) / NB + 1; }
static int64_t gn_partials(int64_t NB, int64_t G) { return (NB * gn_max_chunks(NB) * G + 1) / 2 * 2; }   // even: keeps scale/shift 16-byte aligned

extern "C" size_t fyc_groupnorm_workspace_bytes(int64_t NB, int64_t C, int64_t G) {
  return (size_t)(gn_partials(NB, G) * sizeof(float2) + NB * C * 2 * sizeof(float));
}

template <typename T, int V>
static int32_t groupnorm_impl(const T* x, const float* gamma, const float* beta, T* out, int64_t NB, int64_t R, int C,
                              int G, float eps, int silu, void* ws, cudaStream_t st) {
  float2* partials = (float2*)ws;
  float* scale = (float*)(partials + gn_partials(NB, G));
  float* shift = scale + NB * C;
  const int cvn = C / V;
  const int TX = cvn < 256 ? cvn : 256;
  const int RY = 256 / TX;
  int64_t target = gn_max_chunks(NB) - 1;                                 // CTAs per nb
  int64_t rows_per_cta = ceil_div64(R, target);
  if (rows_per_cta < 4 * RY) rows_per_cta = 4 * RY;
  rows_per_cta = ceil_div64(rows_per_cta, RY) * RY;
  const int chunks = (int)ceil_div64(R, rows_per_cta);
  FYC_CHECK(chunks <= gn_max_chunks(NB), "groupnorm: internal chunk count");
  const size_t smem = (size_t)RY * C * 2 * sizeof(float);
  FYC_CHECK(smem <= 200 * 1024, "groupnorm: C=%d too large", C);
  auto kern = gn_stats_kernel<T, V>;
  if (smem > 48 * 1024) FYC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)chunks, (unsigned)NB);
  kern<<<grid, 256, smem, st>>>(x, partials, R, C, G, rows_per_cta);
  FYC_LAUNCH_CHECK();
  gn_finalize_kernel<<<(unsigned)NB, 256, 2 * G * sizeof(float), st>>>(partials, chunks, gamma, beta, scale, shift, C, G,
                                                                        (double)R * (C / G), eps);
  FYC_LAUNCH_CHECK();
  int64_t total_vec = NB * R * cvn;
  int64_t blocks = ceil_div64(total_vec, 256);
  int64_t cap = (int64_t)fyc_sm_count() * 16;
  unsigned gb = (unsigned)(blocks > cap ? cap : blocks);
  if (silu) gn_apply_kernel<T, V, true><<<gb, 256, 0, st>>>(x, scale, shift, out, R, C, total_vec);
  else gn_apply_kernel<T, V, false><<<gb, 256, 0, st>>>(x, scale, shift, out, R, C, total_vec);
  FYC_LAUNCH_CHECK();
  return FYC_OK;
}

extern "C" int32_t fyc_groupnorm(const void* x, const float* gamma, const float* beta, void* out, int64_t NB, int64_t R,
                                 int64_t C, int64_t G, float eps, int32_t silu, int32_t dtype, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  FYC_CHECK(G > 0 && C % G == 0, "groupnorm: C=%lld not divisible by G=%lld", (long long)C, (long long)G);
  FYC_CHECK(workspace && workspace_bytes >= fyc_groupnorm_workspace_bytes(NB, C, G), "groupnorm: workspace too small");
  FYC_CHECK(NB > 0 && NB < 65536 && R > 0 && C < (1 << 20), "groupnorm: bad shape");
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

extern "C" int32_t fyc_layernorm(const void* x, const float* gamma, const float* beta, void* out, int64_t M, int64_t C,
                                 float eps, const float* pe, int64_t rows_per_frame, int64_t frames, int32_t dtype,
                                 void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  FYC_CHECK(M > 0 && C > 0, "layernorm: bad shape");
  if (pe) FYC_CHECK(rows_per_frame > 0 && frames > 0, "layernorm: pe needs rows_per_frame/frames");
  unsigned grid = (unsigned)ceil_div64(M, 8);
  if (dtype == FYC_BF16) {
    FYC_CHECK(C % 8 == 0 && C <= 8 * 32 * 8, "layernorm(bf16): C=%lld must be a multiple of 8 and <= 2048", (long long)C);
    if (C <= 8 * 32 * 2) layernorm_kernel<bf16, 8, 2><<<grid, 256, 0, st>>>((const bf16*)x, gamma, beta, (bf16*)out, M, (int)C, eps, pe, rows_per_frame, frames);
    else if (C <= 8 * 32 * 5) layernorm_kernel<bf16, 8, 5><<<grid, 256, 0, st>>>((const bf16*)x, gamma, beta, (bf16*)out, M, (int)C, eps, pe, rows_per_frame, frames);
    else layernorm_kernel<bf16, 8, 8><<<grid, 256, 0, st>>>((const bf16*)x, gamma, beta, (bf16*)out, M, (int)C, eps, pe, rows_per_frame, frames);
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
