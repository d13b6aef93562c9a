// Microbenchmark (not part of the product): issue rate / latency of tcgen05.mma (M=128, K=16) chains.
//   umma_bench : for N in {16,32,64,128,256} and nacc in {1,2,4}: cycles per MMA over 256 MMAs that rotate over
//   `nacc` independent TMEM accumulators (nacc = 1: every MMA depends on the previous one's accumulator).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../controllable-latent-diffusion-for-traffic-simulation_b200/csrc/tc_common.cuh"
using namespace cld::tc;

__global__ void __launch_bounds__(128, 1) bench(int N, int nacc, int cnt, int same_ab, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (64 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_base_s), 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t ad0 = make_desc_sw128(smem_u32(smem), 1024);
    const uint64_t bd0 = make_desc_sw128(smem_u32(smem) + 32768, 1024);
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < cnt; ++i) {
        const int a = i % nacc;
        const int kk = same_ab ? 0 : (i & 3);
        umma_bf16(tmem_base + a * N, ad0 + 2 * kk, bd0 + 2 * kk, idesc, i >= nacc ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 1;
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int cnt = 256;
  for (int N : {16, 32, 64, 128, 256})
    for (int nacc : {1, 2, 4}) {
      if (nacc * N > 512) continue;
      long long h[2];
      for (int rep = 0; rep < 2; ++rep) {
        bench<<<grid, 128, 65536>>>(N, nacc, cnt, 0, d);
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      }
      cudaError_t e = cudaGetLastError();
      printf("N=%3d nacc=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.1f) %s\n", N, nacc, (double)h[0] / cnt,
             (double)h[1] / cnt, N * 0.5, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
