// Fused posterior update of the sampler (reference models/dm/dm_model.py:144-163):
//   DDPM: mean = x_t_cof[t]*x - noise_cof[t]*eps ;  x' = mean + 1[t!=0]*exp(.5*logvar[t])*noise
//   DDIM (eta=0; restated from the buffers of dm_model.py:42-43):
//         x0 = sqrt_recip[t]*x - sqrt_recipm1[t]*eps ; x' = sqrt_acp[s]*x0 + sqrt(1-acp[s])*eps
// HBM-bound elementwise work: one float4 per thread per tensor, coalesced; noise either read from
// a caller tensor (parity mode) or generated in-kernel with Philox4x32-10 + Box-Muller
// (throughput mode: saves 1/4 of the traffic).
#include "common.cuh"

namespace cld {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// 4 standard normals for (seed, sequence id, element-quad index)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t seq, uint64_t idx) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)seq, (uint32_t)(seq >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const float two_pow_m32 = 2.3283064365386963e-10f;
  float u0 = ((float)c[0] + 0.5f) * two_pow_m32, u1 = ((float)c[1] + 0.5f) * two_pow_m32;
  float u2 = ((float)c[2] + 0.5f) * two_pow_m32, u3 = ((float)c[3] + 0.5f) * two_pow_m32;
  u0 = fminf(u0, 0.99999994f); u2 = fminf(u2, 0.99999994f);
  float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.f * u1, &s0, &c0);
  sincospif(2.f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

struct StepCoef {
  float a, b;       // mean = a*x - b*eps              (DDPM)
  float c, d;       // mean = c*x0 + d*eps, x0 = a*x - b*eps   (DDIM)
  float sigma;      // 0 when t == 0 or DDIM
  int ddim, final_ddim;
};

__global__ void __launch_bounds__(256) posterior_step_kernel(const float4* __restrict__ x, const float4* __restrict__ eps,
                                                             const float4* __restrict__ noise, uint64_t seed,
                                                             uint64_t seq, uint64_t idx_base, StepCoef k, float4* __restrict__ x_out,
                                                             float4* __restrict__ mean_out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 xv = x[i], ev = eps[i], m;
    m.x = k.a * xv.x - k.b * ev.x; m.y = k.a * xv.y - k.b * ev.y;
    m.z = k.a * xv.z - k.b * ev.z; m.w = k.a * xv.w - k.b * ev.w;
    if (k.ddim && !k.final_ddim) {
      m.x = k.c * m.x + k.d * ev.x; m.y = k.c * m.y + k.d * ev.y;
      m.z = k.c * m.z + k.d * ev.z; m.w = k.c * m.w + k.d * ev.w;
    }
    if (mean_out) mean_out[i] = m;
    if (x_out) {
      float4 o = m;
      if (k.sigma != 0.f) {
        float4 nz = noise ? noise[i] : philox_normal4(seed, seq, idx_base + i);
        o.x += k.sigma * nz.x; o.y += k.sigma * nz.y; o.z += k.sigma * nz.z; o.w += k.sigma * nz.w;
      }
      x_out[i] = o;
    }
  }
}

__global__ void __launch_bounds__(256) add_noise_kernel(const float4* __restrict__ mean, const float4* __restrict__ noise,
                                                        uint64_t seed, uint64_t seq, uint64_t idx_base, float sigma,
                                                        float4* __restrict__ x_out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 o = mean[i];
    if (sigma != 0.f) {
      float4 nz = noise ? noise[i] : philox_normal4(seed, seq, idx_base + i);
      o.x += sigma * nz.x; o.y += sigma * nz.y; o.z += sigma * nz.z; o.w += sigma * nz.w;
    }
    x_out[i] = o;
  }
}

// x_init drawn in-kernel: the same Philox stream family, sequence id ~0 (never a step index)
__global__ void __launch_bounds__(256) philox_fill_kernel(uint64_t seed, uint64_t seq, uint64_t idx_base, float4* __restrict__ out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    out[i] = philox_normal4(seed, seq, idx_base + i);
}

__global__ void fill_t_kernel(int64_t* t, int64_t v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = v;
}

static int grid_for(const CldHandle* h, size_t n4) {
  size_t blocks = (n4 + 255) / 256;
  size_t cap = (size_t)h->num_sms * 8;
  return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

int posterior_step(CldHandle* h, const float* x, const float* eps, const float* noise, uint64_t seed, uint64_t seq,
                   uint64_t idx_base, int t, int t_next, int sampler, float* x_out, float* mean_out, int R, cudaStream_t s) {
  const Schedule& sc = h->sched;
  if (!sc.loaded) return fail(h, CLD_ERR_STATE, "schedule not set");
  if (t < 0 || t >= (int)sc.x_t_cof.size() || t_next >= (int)sc.x_t_cof.size())
    return fail(h, CLD_ERR_ARG, "step index out of range: t=%d t_next=%d", t, t_next);
  StepCoef k{};
  if (sampler == CLD_SAMPLER_DDPM) {
    k.a = sc.x_t_cof[t]; k.b = sc.noise_cof[t];
    k.sigma = (t == 0) ? 0.f : expf(0.5f * sc.logvar[t]);
  } else {
    k.ddim = 1; k.a = sc.sqrt_recip[t]; k.b = sc.sqrt_recipm1[t];
    k.final_ddim = t_next < 0;
    if (t_next >= 0) { k.c = sc.sqrt_acp[t_next]; k.d = sc.sqrt_1macp[t_next]; }
  }
  size_t n = (size_t)R * h->cfg.horizon * h->cfg.latent_dim;
  if (n % 4) return fail(h, CLD_ERR_ARG, "R*T*D must be a multiple of 4");
  posterior_step_kernel<<<grid_for(h, n / 4), 256, 0, s>>>((const float4*)x, (const float4*)eps, (const float4*)noise,
                                                           seed, seq, idx_base, k, (float4*)x_out, (float4*)mean_out, n / 4);
  CLD_LAUNCH_OK(h, "posterior_step_kernel");
  return 0;
}

int add_noise(CldHandle* h, const float* mean, const float* noise, uint64_t seed, uint64_t seq, uint64_t idx_base, int t, float* x_out,
              int R, cudaStream_t s) {
  const Schedule& sc = h->sched;
  if (!sc.loaded) return fail(h, CLD_ERR_STATE, "schedule not set");
  if (t < 0 || t >= (int)sc.logvar.size()) return fail(h, CLD_ERR_ARG, "step index out of range: t=%d", t);
  float sigma = (t == 0) ? 0.f : expf(0.5f * sc.logvar[t]);
  size_t n = (size_t)R * h->cfg.horizon * h->cfg.latent_dim;
  add_noise_kernel<<<grid_for(h, n / 4), 256, 0, s>>>((const float4*)mean, (const float4*)noise, seed, seq, idx_base, sigma,
                                                      (float4*)x_out, n / 4);
  CLD_LAUNCH_OK(h, "add_noise_kernel");
  return 0;
}

int philox_fill(CldHandle* h, uint64_t seed, uint64_t seq, uint64_t idx_base, float* out, int R, cudaStream_t s) {
  size_t n = (size_t)R * h->cfg.horizon * h->cfg.latent_dim;
  philox_fill_kernel<<<grid_for(h, n / 4), 256, 0, s>>>(seed, seq, idx_base, (float4*)out, n / 4);
  CLD_LAUNCH_OK(h, "philox_fill_kernel");
  return 0;
}

int fill_t(CldHandle* h, int64_t* t, int value, int R, cudaStream_t s) {
  fill_t_kernel<<<(R + 255) / 256, 256, 0, s>>>(t, (int64_t)value, R);
  CLD_LAUNCH_OK(h, "fill_t_kernel");
  return 0;
}

}  // namespace cld
