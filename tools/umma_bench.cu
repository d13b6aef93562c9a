// Microbenchmark (not part of the product): execution rate of tcgen05.mma (M=128, K=16, bf16, SS mode: both operands in
// shared memory) as a function of N, with and without a concurrent stream of bulk copies (UBLKCP, global -> shared) into a
// 4 x 16 KB ring, which is what the denoiser's weight producer does.  Answers: is an SS-mode MMA bound by its operand
// fetch from shared memory, and how much of that bandwidth do the weight copies take?
//   umma_bench [grid] [a_off bytes: 0 | 128 .. 896]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../controllable-latent-diffusion-for-traffic-simulation_b200/csrc/tc_common.cuh"
using namespace cld::tc;

// fresh != 0: every MMA reads a different A tile (16 x 4 KB) and a different B tile (rotating over 64 KB), as in the denoiser where
// no operand is re-read by the next instruction; fresh == 0: the same four K slices over and over
__global__ void __launch_bounds__(128, 1) bench(int N, int nacc, int cnt, int copy_bytes, int fresh, const uint8_t* gsrc, long long* out, int a_off) {
  extern __shared__ __align__(1024) uint8_t smem[];     // [0,64K) A, [64K,128K) B, [128K,192K) copy ring
  __shared__ __align__(8) uint64_t bar, cbar[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&cbar[i]), 1);
    fence_barrier_init();
    done = 0;
  }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_base_s), 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t ad0 = make_desc_sw128(smem_u32(smem) + a_off, 1024);     // a_off: A start inside the 1024-byte swizzle atom (row-shifted tile)
    const uint64_t bd0 = make_desc_sw128(smem_u32(smem) + 65536, 1024);
    long long t0 = clock64();
    if (elect_one()) {
      // the loop itself must not be the bottleneck: offsets advance by adds and masks only, 4 K-steps per iteration
      const uint32_t b_step = (uint32_t)(N * 8), b_mask = (uint32_t)(65536 / 16 - 1);
      uint32_t ao = 0, bo = 0;
      for (int i = 0; i < cnt; i += 4) {
        const uint32_t d = tmem_base + (uint32_t)((i >> 2) % nacc) * N;
        const uint64_t a = ad0 + ao, b = bd0 + bo;
        umma_bf16(d, a, b, idesc, i >= 4 * nacc ? 1u : 0u);
        umma_bf16(d, a + 2, b + 2, idesc, 1u);
        umma_bf16(d, a + 4, b + 4, idesc, 1u);
        umma_bf16(d, a + 6, b + 6, idesc, 1u);
        if (fresh) { ao = (ao + 1024u) & 4095u; bo = (bo + b_step) & b_mask; }
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    done = 1;
    if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 1 && copy_bytes > 0) {
    // copy stream: keep 4 bulk copies of copy_bytes in flight until the MMAs are done
    long long copied = 0;
    uint32_t par = 0;
    const uint8_t* src = gsrc + (size_t)blockIdx.x * 65536;
    if (elect_one()) {
      for (int s = 0; s < 4; ++s) {
        mbar_arrive_expect_tx(smem_u32(&cbar[s]), copy_bytes);
        bulk_g2s(smem_u32(smem) + 131072 + s * 16384, src + s * 16384, copy_bytes, smem_u32(&cbar[s]));
      }
      int s = 0;
      while (!done) {
        mbar_wait(smem_u32(&cbar[s]), (par >> s) & 1u);
        par ^= 1u << s;
        copied += copy_bytes;
        mbar_arrive_expect_tx(smem_u32(&cbar[s]), copy_bytes);
        bulk_g2s(smem_u32(smem) + 131072 + s * 16384, src + s * 16384, copy_bytes, smem_u32(&cbar[s]));
        s = (s + 1) & 3;
      }
      for (int k = 0; k < 4; ++k) { mbar_wait(smem_u32(&cbar[s]), (par >> s) & 1u); par ^= 1u << s; s = (s + 1) & 3; }
      if (blockIdx.x == 0) out[2] = copied;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 1;
  const int a_off = argc > 2 ? atoi(argv[2]) : 0;
  long long* d; cudaMalloc(&d, 32);
  uint8_t* src; cudaMalloc(&src, (size_t)grid * 65536); cudaMemset(src, 0, (size_t)grid * 65536);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
  const int cnt = 1024;
  for (int fresh : {0, 1})
  for (int copy_bytes : {0, 16384})
    for (int N : {64, 128, 256})
    for (int nacc : {1, 2}) {
      if (nacc * N > 512) continue;
      long long h[4] = {0, 0, 0, 0};
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(d, 0, 32);
        bench<<<grid, 128, 196608>>>(N, nacc, cnt, copy_bytes, fresh, src, d, a_off);
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      }
      cudaError_t e = cudaGetLastError();
      const double cyc = (double)h[1] / cnt;
      printf("a_off %3d grid %3d fresh %d N=%3d nacc=%d copies %5d B: %.1f cyc/MMA (math floor %.1f), operand bytes/cyc %.1f, copy bytes/cyc %.1f %s\n", a_off, grid, fresh, N, nacc,
             copy_bytes, cyc, N * 0.5, (4096.0 + N * 32.0) / cyc, (double)h[2] / (double)h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
