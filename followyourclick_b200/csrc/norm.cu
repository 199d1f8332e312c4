// GroupNorm (cross-frame or per-frame statistics) and LayerNorm.  Both are HBM-bound: one read for the
// statistics, one read + one write for the apply; statistics are accumulated in fp32 per thread, combined in
// fp64 across CTAs so that E[x^2]-E[x]^2 does not cancel.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// stats: x viewed as [NB, R, C]; grid (chunks, NB); each CTA reduces rows [r0, r1) for all channels.
template <typename T, int V>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, double* __restrict__ sums, int64_t R,
                                                       int C, int G, int64_t rows_per_cta) {
  extern __shared__ float s_acc[];   // [2 * G]
  const int cpg = C / G;
  const int cvn = C / V;
  const int TX = cvn < 256 ? cvn : 256;
  const int RY = 256 / TX;
  const int tx = threadIdx.x % TX, ry = threadIdx.x / TX;
  for (int i = threadIdx.x; i < 2 * G; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int64_t nb = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = (r0 + rows_per_cta < R) ? r0 + rows_per_cta : R;
  const T* base = x + nb * R * C;
  if (ry < RY) {
    for (int cv = tx; cv < cvn; cv += TX) {
      float s[V], q[V];
#pragma unroll
      for (int e = 0; e < V; ++e) { s[e] = 0.f; q[e] = 0.f; }
      for (int64_t r = r0 + ry; r < r1; r += RY) {
        float f[8];
        if constexpr (V == 8) Vec8<T>::load(base + r * C + cv * V, f);
        else if constexpr (V == 4) Vec4<T>::load(base + r * C + cv * V, f);
        else f[0] = to_f(base[r * C + cv]);
#pragma unroll
        for (int e = 0; e < V; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
      }
      // flush: consecutive channels mostly share a group -> merge before the shared atomics
      int g_prev = (cv * V) / cpg;
      float as = 0.f, aq = 0.f;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        int g = (cv * V + e) / cpg;
        if (g != g_prev) {
          atomicAdd(&s_acc[2 * g_prev], as); atomicAdd(&s_acc[2 * g_prev + 1], aq);
          as = 0.f; aq = 0.f; g_prev = g;
        }
        as += s[e]; aq += q[e];
      }
      atomicAdd(&s_acc[2 * g_prev], as); atomicAdd(&s_acc[2 * g_prev + 1], aq);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += 256) atomicAdd(&sums[nb * 2 * G + i], (double)s_acc[i]);
}

// finalize: per (nb, c): scale = rstd * gamma, shift = beta - mean * scale   (same form as ATen's CPU kernel)
__global__ void gn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ scale, float* __restrict__ shift,
                                   int64_t NB, int C, int G, double count, float eps) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= NB * C) return;
  int c = (int)(i % C);
  int64_t nb = i / C;
  int g = c / (C / G);
  double mean = sums[(nb * G + g) * 2] / count;
  double var = sums[(nb * G + g) * 2 + 1] / count - mean * mean;
  if (var < 0) var = 0;
  float rstd = (float)(1.0 / sqrt(var + (double)eps));
  float sc = rstd * gamma[c];
  scale[i] = sc;
  shift[i] = beta[c] - (float)mean * sc;
}

template <typename T, int V, bool SILU>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, T* __restrict__ out, int64_t R,
                                                       int C, int64_t total_vec) {
  const int cvn = C / V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    int cv = (int)(i % cvn);
    int64_t row = i / cvn;
    int64_t nb = row / R;
    const float* sc = scale + nb * C + cv * V;
    const float* sh = shift + nb * C + cv * V;
    float f[8];
    if constexpr (V == 8) Vec8<T>::load(x + i * V, f);
    else if constexpr (V == 4) Vec4<T>::load(x + i * V, f);
    else f[0] = to_f(x[i]);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float y = fmaf(f[e], sc[e], sh[e]);
      f[e] = SILU ? silu_f(y) : y;
    }
    if constexpr (V == 8) Vec8<T>::store(out + i * V, f);
    else if constexpr (V == 4) Vec4<T>::store(out + i * V, f);
    else out[i] = from_f<T>(f[0]);
  }
}

extern "C" size_t fyc_groupnorm_workspace_bytes(int64_t NB, int64_t C, int64_t G) {
  return (size_t)(NB * G * 2 * sizeof(double) + NB * C * 2 * sizeof(float));
}

template <typename T, int V>
static int32_t groupnorm_impl(const T* x, const float* gamma, const float* beta, T* out, int64_t NB, int64_t R, int C,
                              int G, float eps, int silu, void* ws, cudaStream_t st) {
  double* sums = (double*)ws;
  float* scale = (float*)(sums + NB * G * 2);
  float* shift = scale + NB * C;
  FYC_CUDA(cudaMemsetAsync(sums, 0, NB * G * 2 * sizeof(double), st));
  const int cvn = C / V;
  const int TX = cvn < 256 ? cvn : 256;
  const int RY = 256 / TX;
  int64_t target = ((int64_t)fyc_sm_count() * 4 + NB - 1) / NB;          // CTAs per nb
  int64_t rows_per_cta = ceil_div64(R, target);
  if (rows_per_cta < 4 * RY) rows_per_cta = 4 * RY;
  rows_per_cta = ceil_div64(rows_per_cta, RY) * RY;
  dim3 grid((unsigned)ceil_div64(R, rows_per_cta), (unsigned)NB);
  gn_stats_kernel<T, V><<<grid, 256, 2 * G * sizeof(float), st>>>(x, sums, R, C, G, rows_per_cta);
  FYC_LAUNCH_CHECK();
  gn_finalize_kernel<<<(unsigned)ceil_div64(NB * C, 256), 256, 0, st>>>(sums, gamma, beta, scale, shift, NB, C, G,
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
      float o[V];
#pragma unroll
      for (int e = 0; e < V; ++e) {
        int c = cv * V + e;
        float y = (v[i][e] - mean) * rstd * gamma[c] + beta[c];
        o[e] = per ? y + per[c] : y;
      }
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
