// Shared helpers for libfyc_sm100a (B200 / sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fyc.h"

typedef __nv_bfloat16 bf16;

void fyc_set_error(const char* fmt, ...);

#define FYC_CHECK(cond, ...)                 \
  do {                                       \
    if (!(cond)) {                           \
      fyc_set_error(__VA_ARGS__);            \
      return FYC_ERR_INVALID;                \
    }                                        \
  } while (0)

#define FYC_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      fyc_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));      \
      return FYC_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define FYC_LAUNCH_CHECK() FYC_CUDA(cudaGetLastError())

// switch over the storage dtype; body sees `T`
#define FYC_DISPATCH(dt, ...)                                  \
  switch (dt) {                                                \
    case FYC_F32: { using T = float; __VA_ARGS__; } break;     \
    case FYC_BF16: { using T = bf16; __VA_ARGS__; } break;     \
    default: FYC_CHECK(false, "unknown dtype %d", (int)(dt));  \
  }

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8-element vector of T (16 B for bf16, 32 B for fp32) <-> float[8]
template <typename T> struct Vec8;
template <> struct Vec8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float* f) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  static __device__ __forceinline__ void store(bf16* p, const float* f) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float* f) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// 4-element vector
template <typename T> struct Vec4;
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float* f) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
  static __device__ __forceinline__ void store(bf16* p, const float* f) {
    uint2 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
    h[0] = __floats2bfloat162_rn(f[0], f[1]);
    h[1] = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = u;
  }
};
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float* f) {
    float4 a = *reinterpret_cast<const float4*>(p);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
// bf16-storage paths: 2-ulp intrinsics are far below the 2^-9 output rounding
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// exact-erf GELU (F.gelu default; diffusers/models/attention.py:815)
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// exact-erf GELU with erf(z) = 1 - 2^q(z), q = degree-5 least-squares fit of log2(erfc(z)) on [0, 4] (|erf error| <= 7.2e-7,
// |gelu error| <= 1.3e-6 over all x; fit script in DESIGN.md): ONE MUFU.EX2 + 7 FMAs instead of erff()'s ~30 instructions.
// (A first version with Abramowitz-Stegun 7.1.26 needed rcp + ex2 = two MUFU ops per element and made the GEGLU epilogue
// MUFU-bound: 431 vs 636 TFLOP/s on the level-0 FF1 GEMM.)
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float q = fmaf(-0.002980560529977083f, z, 0.02972414717078209f);
  q = fmaf(q, z, -0.14882677793502808f);
  q = fmaf(q, z, -0.9184384942054749f);
  q = fmaf(q, z, -1.6278971433639526f);
  q = fmaf(q, z, -2.8457714051910443e-07f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  const float erfv = copysignf(1.0f - e, x);
  return 0.5f * x * (1.0f + erfv);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
int fyc_sm_count();
