// Micro-benchmark (diagnostic): what a streaming kernel can reach on this GPU for the tensor sizes the norm kernels see.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void copy1(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) b[i] = a[i];
}
template <int U>
__global__ void copyU(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < n; i0 += stride * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i0 + u * stride < n) v[u] = a[i0 + u * stride];
#pragma unroll
    for (int u = 0; u < U; ++u) if (i0 + u * stride < n) b[i0 + u * stride] = v[u];
  }
}
__global__ void readsum(const uint4* __restrict__ a, uint32_t* out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < n; i0 += stride * 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (i0 + u * stride < n) ? a[i0 + u * stride] : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}
__global__ void fill(uint4* b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) b[i] = make_uint4(1, 2, 3, 4);
}
template <typename F> float timeit(F f, int reps = 20) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(e0); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}
int main() {
  for (size_t mb : {21, 84, 336, 1024}) {
    size_t bytes = mb << 20, n = bytes / 16;
    uint4 *a, *b; uint32_t* o; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 4);
    cudaMemset(a, 1, bytes); cudaMemset(b, 0, bytes);
    unsigned g1 = (unsigned)((n + 255) / 256);
    float t1 = timeit([&] { copy1<<<g1, 256>>>(a, b, n); });
    float t4 = timeit([&] { copyU<4><<<148 * 16, 256>>>(a, b, n); });
    float t8 = timeit([&] { copyU<8><<<148 * 8, 256>>>(a, b, n); });
    float tr = timeit([&] { readsum<<<148 * 16, 256>>>(a, o, n); });
    float tw = timeit([&] { fill<<<g1, 256>>>(b, n); });
    float tm = timeit([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); });
    printf("%5zu MB: copy1 %6.1f us %5.2f TB/s | copy4 %5.2f | copy8 %5.2f | memcpy %5.2f TB/s (r+w) | read %5.2f TB/s | write %5.2f TB/s  (%s)\n", mb, t1 * 1e3,
           2 * bytes / t1 / 1e9, 2 * bytes / t4 / 1e9, 2 * bytes / t8 / 1e9, 2 * bytes / tm / 1e9, bytes / tr / 1e9, bytes / tw / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(a); cudaFree(b); cudaFree(o);
  }
  return 0;
}
