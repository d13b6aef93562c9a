// Microbenchmark (not part of the product): FFMA vs FFMA2 issue throughput per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[16]; uint64_t p[16];
  for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((uint64_t)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
  const float w = 1.0001f, z = 0.5f;
  const uint64_t w2 = ((uint64_t)__float_as_uint(w) << 32) | __float_as_uint(w), z2 = ((uint64_t)__float_as_uint(z) << 32) | __float_as_uint(z);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(w), "f"(z));
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) p[i] = fma2(p[i], w2, z2);
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&c, 8);
  for (int warps : {4, 8, 16, 32}) {
    long long h0, h1; const int iters = 2000;
    k<0><<<148, warps * 32>>>(o, iters, c); cudaMemcpy(&h0, c, 8, cudaMemcpyDeviceToHost);
    k<1><<<148, warps * 32>>>(o, iters, c); cudaMemcpy(&h1, c, 8, cudaMemcpyDeviceToHost);
    double n = (double)iters * 16 * warps;   // warp-instructions per SM
    printf("warps/SM %2d: FFMA %.2f warp-inst/clk/SM (%.1f FMA/clk) | FFMA2 %.2f warp-inst/clk/SM (%.1f FMA/clk)\n", warps, n / h0, n / h0 * 32, n / h1, n / h1 * 64);
  }
  return 0;
}
