// Microbenchmark (not part of the product): throughput of tcgen05.ld (tensor memory -> registers) for the shapes the denoiser's
// epilogue could use: 32x32b.x8 / .x16 / .x32, with 4, 8 or 16 warps reading (warp & 3 = lane quadrant), one wait per `per_wait` loads.
//   tmem_ld_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../controllable-latent-diffusion-for-traffic-simulation_b200/csrc/tc_common.cuh"
using namespace cld::tc;

template <int X>
__global__ void __launch_bounds__(512, 1) bench(int iters, int per_wait, long long* out) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_base_s), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i += per_wait) {
    for (int j = 0; j < per_wait; ++j) {
      const uint32_t col = (uint32_t)(((i + j) * X) & 511 & ~(X - 1));
      if (X == 8) { uint32_t r[8]; tmem_ld8(base + col, r); acc ^= r[0] ^ r[7]; }
      else if (X == 16) { uint32_t r[16]; tmem_ld16(base + col, r); acc ^= r[0] ^ r[15]; }
      else { uint32_t r[32]; tmem_ld32(base + col, r); acc ^= r[0] ^ r[31]; }
    }
    tmem_wait_ld();
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[3] = acc;
  __syncthreads();
  if (tid == 0) { out[0] = t1 - t0; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int iters = 1024;
  for (int warps : {4, 8, 16})
    for (int x : {8, 16, 32})
      for (int per_wait : {1, 4}) {
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) {
          if (x == 8) bench<8><<<1, warps * 32>>>(iters, per_wait, d);
          else if (x == 16) bench<16><<<1, warps * 32>>>(iters, per_wait, d);
          else bench<32><<<1, warps * 32>>>(iters, per_wait, d);
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        }
        const double bytes = (double)warps * iters * 32 * x * 4;
        printf("warps %2d x%-2d per_wait %d: %.1f cycles per load per warp, %.1f B/cycle per SM %s\n", warps, x, per_wait, (double)h[0] / iters,
               bytes / (double)h[0], cudaGetLastError() == cudaSuccess ? "" : cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
