// Micro-benchmark (diagnostic): per-SM throughput of the exp2 variants and packed fp32 ops the attention softmax could use.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t b0 = __float_as_uint(a0) >> 3, b1 = b0 + 7, b2 = b0 + 11, b3 = b0 + 13;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {   // ex2.approx.ftz.f32: 4 independent chains
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {   // ex2.approx.ftz.bf16x2
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b0)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b2)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b3));
    } else if (MODE == 2) {   // fma.rn.f32x2
      uint64_t p0, p1; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1)); asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p0) : "l"(p1)); asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(p1) : "l"(p0));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p0) : "l"(p1)); asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(p1) : "l"(p0));
      asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0)); asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
    } else if (MODE == 3) {   // plain FFMA x4
      a0 = fmaf(a0, a1, a2); a1 = fmaf(a1, a2, a3); a2 = fmaf(a2, a3, a0); a3 = fmaf(a3, a0, a1);
    } else if (MODE == 4) {   // cvt.rn.bf16x2.f32 (pack) x4
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b0) : "f"(a0), "f"(a1)); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b1) : "f"(a1), "f"(a2));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b2) : "f"(a2), "f"(a3)); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b3) : "f"(a3), "f"(a0));
      a0 += __uint_as_float(b0); a1 += __uint_as_float(b1);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(b0 ^ b1 ^ b2 ^ b3);
}
template <int MODE> void run(const char* name, int per_iter_ops) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000;
  k<MODE><<<148 * 2, 256>>>(out, 100, 0.5f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<148 * 4, 512>>>(out, iters, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = 148.0 * 4 * 16 * iters * per_iter_ops;   // warp-level instructions
  printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM @1.9GHz  (err %s)\n", name, ms, warp_instr / (ms * 1e-3) / 148 / 1.9e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  run<0>("ex2.approx.ftz.f32", 4); run<1>("ex2.approx.ftz.bf16x2", 4); run<2>("fma.rn.f32x2", 4); run<3>("fma.rn.f32", 4); run<4>("cvt.rn.bf16x2.f32 (+2 fadd)", 6);
  return 0;
}
