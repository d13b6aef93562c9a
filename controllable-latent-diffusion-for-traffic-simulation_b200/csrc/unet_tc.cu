// bf16 tcgen05 denoiser path -- placeholder until the tensor-core kernel lands.
#include "unet_tc.cuh"

namespace cld {
bool tc_enabled(const CldHandle* h) { return false && h->cfg.precision == CLD_PREC_BF16; }
int tc_pack_block(CldHandle*, int, int, int, const float*, const float*, const float*, const float*, const float*,
                  const float*, const float*, const float*, const float*, const float*, cudaStream_t) { return 0; }
int tc_pack_down(CldHandle*, int, int, const float*, const float*, cudaStream_t) { return 0; }
int tc_pack_up(CldHandle*, int, int, const float*, const float*, cudaStream_t) { return 0; }
int tc_pack_final(CldHandle*, const float*, const float*, const float*, const float*, const float*, const float*,
                  cudaStream_t) { return 0; }
int tc_finalize(CldHandle*, cudaStream_t) { return 0; }
int tc_unet_forward(CldHandle* h, const float*, const float*, const int64_t*, float*, int, cudaStream_t) {
  return fail(h, CLD_ERR_UNSUPPORTED, "bf16 tensor-core denoiser not built");
}
void tc_destroy(CldHandle*) {}
}  // namespace cld
