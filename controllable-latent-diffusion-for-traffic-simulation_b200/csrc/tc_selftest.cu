// Hardware self-test of the UMMA building blocks used by unet_tc.cu: one CTA stages a 128B-swizzled
// A image and B image in shared memory, issues `nk16` tcgen05.mma (M=128, K=16 each) with an A descriptor
// that starts `a_start_off` bytes into the image and steps `sbo_a` bytes per 8-row group, and returns
// the fp32 accumulator [128][N].  tests/test_gpu_tc.py compares it with numpy for the exact descriptor
// patterns the denoiser uses (tap shifts by whole 1024-byte swizzle atoms, stride-2 row groups).
#include "common.cuh"
#include "tc_common.cuh"

namespace cld {
using namespace tc;

__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const uint4* __restrict__ a_img, int a_bytes,
                                                             const uint4* __restrict__ b_img, int b_bytes,
                                                             int a_start_off, int sbo_a, int N, int nk16,
                                                             int base_offset, float* __restrict__ d_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 1023) / 1024) * 1024;
  for (int i = tid; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = a_img[i];
  for (int i = tid; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = b_img[i];
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_base_s), ncols); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    for (int k = 0; k < nk16; ++k) {
      uint64_t ad = make_desc_sw128(smem_u32(sa) + a_start_off + k * 32, sbo_a) | ((uint64_t)(base_offset & 7) << 49);
      uint64_t bd = make_desc_sw128(smem_u32(sb) + k * 32, 1024);
      umma_bf16(tmem_base, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
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
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

}  // namespace cld

extern "C" int cld_tc_selftest(const void* a_img, int a_bytes, const void* b_img, int b_bytes, int a_start_off,
                               int sbo_a, int N, int nk16, int base_offset, float* d_out, void* stream) {
  using namespace cld;
  if (!a_img || !b_img || !d_out) return fail(nullptr, CLD_ERR_ARG, "null argument");
  if (a_bytes % 16 || b_bytes % 16 || N % 16 || N < 16 || N > 256 || nk16 < 1 || nk16 > 4)
    return fail(nullptr, CLD_ERR_ARG, "bad selftest shape");
  size_t smem = (size_t)((a_bytes + 1023) / 1024) * 1024 + b_bytes + 1024;
  if (smem > 200 * 1024) return fail(nullptr, CLD_ERR_ARG, "selftest images too large");
  cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(nullptr, CLD_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint4*)a_img, a_bytes, (const uint4*)b_img, b_bytes,
                                                            a_start_off, sbo_a, N, nk16, base_offset, d_out);
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(nullptr, CLD_ERR_CUDA, "tc_selftest launch: %s", cudaGetErrorString(e));
  return CLD_OK;
}
