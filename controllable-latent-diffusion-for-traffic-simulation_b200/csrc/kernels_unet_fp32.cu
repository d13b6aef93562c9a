// fp32 (CUDA-core) denoiser path: the 1e-4-parity mode of TemporalMapUnet.forward
// (reference src/tbsim/models/temporal.py:122-180, diffuser_helpers.py:20-67).
// Activations are channels-last [R, T', C] fp32 so that the latent [R,T,4] needs no rearrange.
// Every convolution (k5 / k1 / k3-stride-2 / the two phases of the k4-stride-2 transposed conv) is
// one implicit GEMM:  out[r, j*ostride+ooff, :] = bias + sum_tap in[r, j*istride+ioff[tap], :] @ W[tap]
#include "common.cuh"

namespace cld {

struct ConvArgs {
  const float* in0; int c0;     // first source  [R, Tin, c0]
  const float* in1; int c1;     // optional concatenated source [R, Tin, c1]
  int Tin;
  const float* w;               // [ntaps][cin][cout]
  const float* bias;            // [cout] or nullptr
  float* out; int Tout;         // out is [R, Tout, cout]
  int cout;
  int ntaps; int ioff[5]; int istride, ostride, ooff;
  int Tj;                       // output positions produced per row by this launch
  int R;
};

constexpr int BM = 128, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) conv_gemm_fp32(ConvArgs a) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int M = a.R * a.Tj;
  const int cin = a.c0 + a.c1;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // A-load coordinates of this thread (2 float4 per stage)
  int a_row[2], a_q[2], a_r[2], a_j[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int idx = tid + i * 256;
    a_row[i] = idx >> 2; a_q[i] = idx & 3;
    int m = m0 + a_row[i];
    a_r[i] = (m < M) ? m / a.Tj : -1;
    a_j[i] = (m < M) ? m % a.Tj : 0;
  }
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;
  const int kchunks = (cin + BK - 1) / BK;
  for (int tap = 0; tap < a.ntaps; ++tap) {
    for (int kc = 0; kc < kchunks; ++kc) {
      const int ci0 = kc * BK;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int c = ci0 + a_q[i] * 4;
        int ti = a_j[i] * a.istride + a.ioff[tap];
        if (a_r[i] >= 0 && c < cin && ti >= 0 && ti < a.Tin) {
          const float* p = (c < a.c0)
              ? a.in0 + ((size_t)a_r[i] * a.Tin + ti) * a.c0 + c
              : a.in1 + ((size_t)a_r[i] * a.Tin + ti) * a.c1 + (c - a.c0);
          v = *reinterpret_cast<const float4*>(p);
        }
        As[a_q[i] * 4 + 0][a_row[i]] = v.x; As[a_q[i] * 4 + 1][a_row[i]] = v.y;
        As[a_q[i] * 4 + 2][a_row[i]] = v.z; As[a_q[i] * 4 + 3][a_row[i]] = v.w;
      }
      {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int c = ci0 + b_k, n = n0 + b_n;
        if (c < cin && n < a.cout)
          v = *reinterpret_cast<const float4*>(a.w + ((size_t)tap * cin + c) * a.cout + n);
        *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  const int n = n0 + tx * 4;
  if (n < a.cout) {
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.bias) bb = *reinterpret_cast<const float4*>(a.bias + n);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + ty * 8 + i;
      if (m < M) {
        int r = m / a.Tj, j = m % a.Tj;
        float4 o = make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
        *reinterpret_cast<float4*>(a.out + ((size_t)r * a.Tout + j * a.ostride + a.ooff) * a.cout + n) = o;
      }
    }
  }
}

__device__ __forceinline__ float mish_f(float x) {
  // nn.Mish: x * tanh(softplus(x)), softplus threshold 20 as in torch
  float sp = (x > 20.f) ? x : log1pf(expf(x));
  return x * tanhf(sp);
}

// GroupNorm(8 groups, eps 1e-5, affine) -> Mish -> (+ per-row channel bias | + residual tensor).
// One CTA per row, warp g owns group g.   (diffuser_helpers.py:58-64; temporal.py:37-45)
__global__ void __launch_bounds__(256) gn_mish_fp32(const float* __restrict__ in, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, const float* __restrict__ tbias,
                                                    int tb_stride, const float* __restrict__ res,
                                                    float* __restrict__ out, int T, int C) {
  const int r = blockIdx.x, g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cpg = C >> 3, n = T * cpg;
  const float* xin = in + (size_t)r * T * C + g * cpg;
  float s = 0.f;
  for (int e = lane; e < n; e += 32) s += xin[(e / cpg) * C + (e % cpg)];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)n;
  float v = 0.f;
  for (int e = lane; e < n; e += 32) {
    float d = xin[(e / cpg) * C + (e % cpg)] - mean;
    v = fmaf(d, d, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = 1.0f / sqrtf(v / (float)n + 1e-5f);
  for (int e = lane; e < n; e += 32) {
    int t = e / cpg, c = g * cpg + (e % cpg);
    float y = (xin[t * C + (e % cpg)] - mean) * rstd * gamma[c] + beta[c];
    y = mish_f(y);
    if (tbias) y += tbias[(size_t)r * tb_stride + c];
    if (res) y += res[((size_t)r * T + t) * C + c];
    out[((size_t)r * T + t) * C + c] = y;
  }
}

// Sinusoidal embedding -> Linear(d,4d) -> Mish -> Linear(4d,d), then Mish over [t_emb, cond]
// (temporal.py:74-79,139-142 and the leading nn.Mish of every block's time_mlp, temporal.py:21-25).
__global__ void __launch_bounds__(128) time_cond_mish(const int64_t* __restrict__ t, const float* __restrict__ cond,
                                                      const float* __restrict__ w1, const float* __restrict__ b1,
                                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                                      const float* __restrict__ freqs, float* __restrict__ tcm,
                                                      int d, int cond_dim) {
  __shared__ float emb[64];
  __shared__ float hid[256];
  const int r = blockIdx.x, tid = threadIdx.x;
  const float tv = (float)t[r];
  const int half = d >> 1;
  if (tid < d) {
    float a = tv * freqs[tid % half];
    emb[tid] = (tid < half) ? sinf(a) : cosf(a);
  }
  __syncthreads();
  for (int o = tid; o < 4 * d; o += blockDim.x) {
    float acc = b1[o];
    for (int k = 0; k < d; ++k) acc = fmaf(w1[o * d + k], emb[k], acc);
    hid[o] = mish_f(acc);
  }
  __syncthreads();
  float* orow = tcm + (size_t)r * (d + cond_dim);
  if (tid < d) {
    float acc = b2[tid];
    for (int k = 0; k < 4 * d; ++k) acc = fmaf(w2[tid * 4 * d + k], hid[k], acc);
    orow[tid] = mish_f(acc);
  }
  for (int c = tid; c < cond_dim; c += blockDim.x) orow[d + c] = mish_f(cond[(size_t)r * cond_dim + c]);
}

int gn_mish_launch(CldHandle* h, const float* in, const GnW& n, const float* tbias, int tb_stride, const float* res, float* out,
                   int T, int C, int R, cudaStream_t s) {
  gn_mish_fp32<<<R, 256, 0, s>>>(in, n.g, n.b, tbias, tb_stride, res, out, T, C);
  CLD_LAUNCH_OK(h, "gn_mish_fp32");
  return 0;
}

static int launch_conv(CldHandle* h, const ConvW& w, const float* in0, int c0, const float* in1, int c1,
                       int Tin, float* out, int Tout, int Tj, int istride, int ostride, int ooff,
                       const int* ioff, const float* bias, int R, cudaStream_t s) {
  ConvArgs a;
  a.in0 = in0; a.c0 = c0; a.in1 = in1; a.c1 = c1; a.Tin = Tin;
  a.w = w.w; a.bias = bias; a.out = out; a.Tout = Tout; a.cout = w.cout;
  a.ntaps = w.ntaps;
  for (int i = 0; i < 5; ++i) a.ioff[i] = (i < w.ntaps) ? ioff[i] : 0;
  a.istride = istride; a.ostride = ostride; a.ooff = ooff; a.Tj = Tj; a.R = R;
  dim3 grid((R * Tj + BM - 1) / BM, (w.cout + BN - 1) / BN);
  conv_gemm_fp32<<<grid, 256, 0, s>>>(a);
  CLD_LAUNCH_OK(h, "conv_gemm_fp32");
  return 0;
}

static const int kOff5[5] = {-2, -1, 0, 1, 2};
static const int kOff3[5] = {-1, 0, 1, 0, 0};
static const int kOff1[5] = {0, 0, 0, 0, 0};

static int run_resblock(CldHandle* h, const ResBlockW& rb, const float* in0, int c0, const float* in1, int c1,
                        int T, float* tmpA, float* tmpB, float* tmpR, float* out, int R, cudaStream_t s) {
  int rc;
  // blocks[0]: conv k5 -> GN -> Mish, + time/cond bias
  if ((rc = launch_conv(h, rb.c0, in0, c0, in1, c1, T, tmpA, T, T, 1, 1, 0, kOff5, rb.c0.b, R, s))) return rc;
  gn_mish_fp32<<<R, 256, 0, s>>>(tmpA, rb.n0.g, rb.n0.b, h->tbias + rb.tb_off, h->unet.tb_total, nullptr, tmpB, T,
                                 rb.cout);
  CLD_LAUNCH_OK(h, "gn_mish_fp32");
  // blocks[1]
  if ((rc = launch_conv(h, rb.c1, tmpB, rb.cout, nullptr, 0, T, tmpA, T, T, 1, 1, 0, kOff5, rb.c1.b, R, s))) return rc;
  const float* res = in0;   // identity residual (only when there is no concat and cin == cout)
  if (rb.res.w) {
    if ((rc = launch_conv(h, rb.res, in0, c0, in1, c1, T, tmpR, T, T, 1, 1, 0, kOff1, rb.res.b, R, s))) return rc;
    res = tmpR;
  }
  gn_mish_fp32<<<R, 256, 0, s>>>(tmpA, rb.n1.g, rb.n1.b, nullptr, 0, res, out, T, rb.cout);
  CLD_LAUNCH_OK(h, "gn_mish_fp32");
  return 0;
}

int unet_stage_elems(const CldHandle* h, int stage) {
  const int T = h->cfg.horizon;
  const int* d = h->cfg.dims;
  switch (stage) {
    case 0: case 1: return T * d[0];
    case 2: return (T / 2) * d[0];
    case 3: case 4: return (T / 2) * d[1];
    case 5: return (T / 4) * d[1];
    case 6: case 7: case 8: case 9: return (T / 4) * d[2];
    case 10: case 11: return (T / 4) * d[1];
    case 12: return (T / 2) * d[1];
    case 13: case 14: return (T / 2) * d[0];
    case 15: case 16: return T * d[0];
    default: return -1;
  }
}

static int tap(CldHandle* h, int stage, const float* buf, int R, cudaStream_t s) {
  if (h->dbg_stage == stage && h->dbg_out) {
    size_t n = (size_t)R * unet_stage_elems(h, stage);
    CLD_CUDA_OK(h, cudaMemcpyAsync(h->dbg_out, buf, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

// time / cond projections of all 12 blocks at once: tbias[R, tb_total] = Mish([t_emb, cond]) @ Wtb + btb
int unet_time_cond(CldHandle* h, const float* cond, const int64_t* t, int R, cudaStream_t s) {
  const UnetW& u = h->unet;
  const CldConfig& c = h->cfg;
  train_invalidate(h);
  time_cond_mish<<<R, 128, 0, s>>>(t, cond, u.t1_w, u.t1_b, u.t2_w, u.t2_b, u.freqs, h->tcm, c.base_dim, c.cond_dim);
  CLD_LAUNCH_OK(h, "time_cond_mish");
  return 0;
}

int unet_time_bias(CldHandle* h, const float* cond, const int64_t* t, int R, cudaStream_t s) {
  const UnetW& u = h->unet;
  const CldConfig& c = h->cfg;
  int rc0 = unet_time_cond(h, cond, t, R, s);
  if (rc0) return rc0;
  ConvW tb; tb.w = u.tb_w; tb.b = u.tb_b; tb.cin = c.base_dim + c.cond_dim; tb.cout = u.tb_total; tb.ntaps = 1;
  return launch_conv(h, tb, h->tcm, tb.cin, nullptr, 0, 1, h->tbias, 1, 1, 1, 1, 0, kOff1, tb.b, R, s);
}

// ---- split form used by the sampler (every row shares the same step index t):
//   tbias[r, c] = cond_bias[r, c] + time_vec[c]
//   cond_bias = Mish(cond) @ Wtb[d:, :] + btb      (step-invariant: computed once per cld_sample call)
//   time_vec  = Mish(t_emb(t)) @ Wtb[:d, :]        (one vector per step, shared by all rows)
// Linear(Mish([t_emb, cond])) of temporal.py:21-25 splits exactly because Mish is elementwise.
__global__ void __launch_bounds__(256) cond_mish_kernel(const float* __restrict__ cond, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = mish_f(cond[i]);
}

__global__ void __launch_bounds__(256) time_vec_kernel(int t, const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                                       const float* __restrict__ freqs, const float* __restrict__ tb_w,
                                                       float* __restrict__ tvec, int d, int tb_total) {
  __shared__ float emb[64];
  __shared__ float hid[256];
  __shared__ float tm[64];
  const int tid = threadIdx.x;
  const float tv = (float)t;
  const int half = d >> 1;
  if (tid < d) {
    float a = tv * freqs[tid % half];
    emb[tid] = (tid < half) ? sinf(a) : cosf(a);
  }
  __syncthreads();
  for (int o = tid; o < 4 * d; o += blockDim.x) {
    float acc = b1[o];
    for (int k = 0; k < d; ++k) acc = fmaf(w1[o * d + k], emb[k], acc);
    hid[o] = mish_f(acc);
  }
  __syncthreads();
  if (tid < d) {
    float acc = b2[tid];
    for (int k = 0; k < 4 * d; ++k) acc = fmaf(w2[tid * 4 * d + k], hid[k], acc);
    tm[tid] = mish_f(acc);
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + tid;
  if (c < tb_total) {
    float acc = 0.f;
    for (int k = 0; k < d; ++k) acc = fmaf(tb_w[(size_t)k * tb_total + c], tm[k], acc);
    tvec[c] = acc;
  }
}

int unet_cond_bias(CldHandle* h, const float* cond, int R, cudaStream_t s) {
  const UnetW& u = h->unet;
  const CldConfig& c = h->cfg;
  const size_t n = (size_t)R * c.cond_dim;
  train_invalidate(h);
  cond_mish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cond, h->tcm, n);
  CLD_LAUNCH_OK(h, "cond_mish_kernel");
  ConvW tb; tb.w = u.tb_w + (size_t)c.base_dim * u.tb_total; tb.b = u.tb_b; tb.cin = c.cond_dim; tb.cout = u.tb_total; tb.ntaps = 1;
  return launch_conv(h, tb, h->tcm, tb.cin, nullptr, 0, 1, h->tbias, 1, 1, 1, 1, 0, kOff1, tb.b, R, s);
}

int unet_time_vec_to(CldHandle* h, int t, float* dst, cudaStream_t s) {
  const UnetW& u = h->unet;
  const CldConfig& c = h->cfg;
  time_vec_kernel<<<(u.tb_total + 255) / 256, 256, 0, s>>>(t, u.t1_w, u.t1_b, u.t2_w, u.t2_b, u.freqs, u.tb_w, dst, c.base_dim, u.tb_total);
  CLD_LAUNCH_OK(h, "time_vec_kernel");
  return 0;
}
int unet_time_vec(CldHandle* h, int t, cudaStream_t s) { return unet_time_vec_to(h, t, h->tvec, s); }

int unet_forward_fp32(CldHandle* h, const float* x, const float* cond, const int64_t* t, float* eps, int R,
                      cudaStream_t s) {
  const UnetW& u = h->unet;
  const CldConfig& c = h->cfg;
  const int T = c.horizon, T2 = T / 2, T4 = T / 4;
  const int d0 = c.dims[0], d1 = c.dims[1], d2 = c.dims[2], D = c.latent_dim;
  float *a0 = h->act[0], *a1 = h->act[1], *sk1 = h->act[2], *sk2 = h->act[3];
  float *tA = h->act[4], *tB = h->act[5], *tR = h->act[6];
  int rc;
  if ((rc = unet_time_bias(h, cond, t, R, s))) return rc;
#define RB(i, in0, c0, in1, c1, TT, out, stage)                                                     \
  if ((rc = run_resblock(h, u.rb[i], in0, c0, in1, c1, TT, tA, tB, tR, out, R, s))) return rc;       \
  if ((rc = tap(h, stage, out, R, s))) return rc;
  RB(0, x, D, nullptr, 0, T, a0, 0)
  RB(1, a0, d0, nullptr, 0, T, a1, 1)
  if ((rc = launch_conv(h, u.down[0], a1, d0, nullptr, 0, T, a0, T2, T2, 2, 1, 0, kOff3, u.down[0].b, R, s))) return rc;
  if ((rc = tap(h, 2, a0, R, s))) return rc;
  RB(2, a0, d0, nullptr, 0, T2, a1, 3)
  RB(3, a1, d1, nullptr, 0, T2, sk1, 4)
  if ((rc = launch_conv(h, u.down[1], sk1, d1, nullptr, 0, T2, a0, T4, T4, 2, 1, 0, kOff3, u.down[1].b, R, s))) return rc;
  if ((rc = tap(h, 5, a0, R, s))) return rc;
  RB(4, a0, d1, nullptr, 0, T4, a1, 6)
  RB(5, a1, d2, nullptr, 0, T4, sk2, 7)
  RB(6, sk2, d2, nullptr, 0, T4, a0, 8)
  RB(7, a0, d2, nullptr, 0, T4, a1, 9)
  RB(8, a1, d2, sk2, d2, T4, a0, 10)
  RB(9, a0, d1, nullptr, 0, T4, a1, 11)
  {
    const int off_e[5] = {0, -1, 0, 0, 0}, off_o[5] = {1, 0, 0, 0, 0};
    if ((rc = launch_conv(h, u.up[0][0], a1, d1, nullptr, 0, T4, a0, T2, T4, 1, 2, 0, off_e, u.up_b[0], R, s))) return rc;
    if ((rc = launch_conv(h, u.up[0][1], a1, d1, nullptr, 0, T4, a0, T2, T4, 1, 2, 1, off_o, u.up_b[0], R, s))) return rc;
    if ((rc = tap(h, 12, a0, R, s))) return rc;
  }
  RB(10, a0, d1, sk1, d1, T2, a1, 13)
  RB(11, a1, d0, nullptr, 0, T2, a0, 14)
  {
    const int off_e[5] = {0, -1, 0, 0, 0}, off_o[5] = {1, 0, 0, 0, 0};
    if ((rc = launch_conv(h, u.up[1][0], a0, d0, nullptr, 0, T2, a1, T, T2, 1, 2, 0, off_e, u.up_b[1], R, s))) return rc;
    if ((rc = launch_conv(h, u.up[1][1], a0, d0, nullptr, 0, T2, a1, T, T2, 1, 2, 1, off_o, u.up_b[1], R, s))) return rc;
    if ((rc = tap(h, 15, a1, R, s))) return rc;
  }
#undef RB
  if ((rc = launch_conv(h, u.fin0, a1, d0, nullptr, 0, T, tA, T, T, 1, 1, 0, kOff5, u.fin0.b, R, s))) return rc;
  gn_mish_fp32<<<R, 256, 0, s>>>(tA, u.fin0n.g, u.fin0n.b, nullptr, 0, nullptr, a0, T, d0);
  CLD_LAUNCH_OK(h, "gn_mish_fp32");
  if ((rc = tap(h, 16, a0, R, s))) return rc;
  if ((rc = launch_conv(h, u.fin1, a0, d0, nullptr, 0, T, eps, T, T, 1, 1, 0, kOff1, u.fin1.b, R, s))) return rc;
  return 0;
}

}  // namespace cld
