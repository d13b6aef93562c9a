// Microbenchmark + correctness probe (not part of the product): tcgen05.mma with the A operand in TENSOR MEMORY
// (M = 128, K = 16 per instruction, fp16 packed two per 32-bit column, lane = row) and B from shared memory
// (K-major, 128B swizzle).  Prints max |D - A B^T| against a host reference and cycles per MMA for N in {16, 32, 64}.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../controllable-latent-diffusion-for-traffic-simulation_b200/csrc/tc_common.cuh"
using namespace cld::tc;

__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A[m][k] = ((m * 7 + k * 3) % 17 - 8) / 8 ; B[n][k] = ((n * 5 + k) % 13 - 6) / 4   (exact in fp16)
__host__ __device__ inline float a_val(int m, int k) { return (float)((m * 7 + k * 3) % 17 - 8) / 8.f; }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 5 + k) % 13 - 6) / 4.f; }

__global__ void __launch_bounds__(128, 1) bench(int N, int cnt, float* d_out, long long* t_out, int mn_major, int sbo_bytes, int kstep_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 64; i += 128) {                 // B tile: up to 64 rows x 64 k
    const int n = i >> 6, k = i & 63;
    const __half hv = __float2half_rn(n < N ? b_val(n, k) : 0.f);
    if (!mn_major) *reinterpret_cast<__half*>(smem + sw128_off(n, k >> 3) + (k & 7) * 2) = hv;
    else if (n < 32) {
      // N-major, 64-byte rows (32 n per k), SWIZZLE_64B: 16-byte chunk index ^= (k >> 1) & 3 ; 8 k per 512-byte atom
      const int chunk = (n >> 3) ^ ((k >> 1) & 3);
      *reinterpret_cast<__half*>(smem + k * 64 + chunk * 16 + (n & 7) * 2) = hv;
    }
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_base_s), 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // A -> TMEM columns [128, 160): lane m = tid, column c holds (A[m][2c], A[m][2c+1]) as fp16x2 (low half = even k)
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t r[8];
    for (int c = 0; c < 8; ++c) {
      __half2 h = __floats2half2_rn(a_val(tid, 2 * (c0 + c)), a_val(tid, 2 * (c0 + c) + 1));
      r[c] = *reinterpret_cast<uint32_t*>(&h);
    }
    tmem_st8(tmem_base + ((uint32_t)(warp * 32) << 16) + 128 + c0, r);
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) | (mn_major ? (1u << 16) : 0u);
  uint64_t bd0 = make_desc_sw128(smem_u32(smem), 1024);
  int kstep = 2;
  if (mn_major) {
    bd0 = (uint64_t)((smem_u32(smem) & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
    kstep = kstep_bytes >> 4;
  }
  if (warp == 0) {
    // correctness: 4 K = 16 steps
    if (elect_one()) {
      for (int k = 0; k < 4; ++k) umma_ts_f16(tmem_base, tmem_base + 128 + 8 * k, bd0 + kstep * k, idesc, k ? 1u : 0u);
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) d_out[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < cnt; ++i) umma_ts_f16(tmem_base + 32 * (i & 1), tmem_base + 128 + 8 * (i & 3), bd0 + kstep * (i & 3), idesc, i >= 2 ? 1u : 0u);
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 1);
    long long t2 = clock64();
    if (tid == 0) { t_out[0] = t1 - t0; t_out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

int main() {
  float* d; long long* t;
  cudaMalloc(&d, 128 * 64 * sizeof(float)); cudaMalloc(&t, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  struct Cfg { int N, mn, sbo, kstep; };
  const Cfg cfgs[] = {{16, 0, 0, 0}, {32, 0, 0, 0}, {64, 0, 0, 0}, {32, 1, 512, 1024}, {32, 1, 1024, 1024}, {32, 1, 512, 512}, {32, 1, 64, 1024}};
  for (const Cfg& cf : cfgs) {
    const int N = cf.N;
    const int cnt = 256;
    bench<<<1, 128, 65536>>>(N, cnt, d, t, cf.mn, cf.sbo, cf.kstep);
    cudaError_t e = cudaDeviceSynchronize();
    float h[128 * 64]; long long ht[2];
    cudaMemcpy(h, d, 128 * N * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(ht, t, 16, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < 64; ++k) ref += (double)a_val(m, k) * b_val(n, k);
        double err = fabs(ref - h[m * N + n]);
        if (err > maxerr) maxerr = err;
        if (fabs(ref) > maxref) maxref = fabs(ref);
      }
    printf("TS-mode N=%2d mn_major=%d sbo=%d kstep=%d: max |D - ref| = %.3e (max |ref| %.1f) | issue %.1f cyc/MMA, complete %.1f cyc/MMA  %s\n", N, cf.mn, cf.sbo, cf.kstep, maxerr, maxref,
           (double)ht[0] / cnt, (double)ht[1] / cnt, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
