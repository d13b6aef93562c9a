// Tensor-core (tcgen05, kind::tf32) implicit-GEMM convolution for the denoiser TRAINING step (reference
// src/tbsim/models/temporal.py:122-180 forward and its autograd data gradient): the stride-1 convolutions of the residual blocks,
// forward and data gradient, on fp32 channels-last activations [R, T', C] WITHOUT any conversion or im2col pass:
//   * the A operand of filter tap `k` and 32-channel block `c` is ONE 3-D TMA box {32 channels, T' slots, rbox rows} of the activation
//     tensor whose time coordinate starts at the tap's offset: the TMA unit zero-fills the slots outside [0, T') (= the padding) and
//     writes the 128B-swizzled K-major tile the MMA descriptor expects.  GEMM row m = row_in_box * T' + t; rbox * T' <= 128.
//   * the B operand is a box {32, BN, 1} of the weights viewed as {K, N, tap}: the forward reads a [tap][cout][cin] copy, the data
//     gradient reads the packed forward weights [tap][cin][cout] as they are (their N is cin, their K is cout).
//   * kind::tf32 takes the fp32 bits from shared memory (10-bit mantissa products, fp32 accumulation in tensor memory): 4 MMAs of
//     K = 8 per 32-channel block.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (owns tensor memory), warps 2-5 epilogue (tensor memory -> registers -> + bias
// (+ previous value) -> global).  One output tile (<= 128 GEMM rows x BN channels) per CTA.
#include <cuda.h>

#include <map>
#include <tuple>

#include "common.cuh"
#include "tc_common.cuh"

namespace cld {
using namespace cld::tc;

constexpr int TF_THREADS = 192;
constexpr int TF_MAX_STAGES = 8;
constexpr int TF_A_BYTES = 128 * 128;            // 128 GEMM rows x 32 fp32
constexpr int TF_MAX_BN = 256;
constexpr int TF_DATA_BYTES = 192 * 1024;        // stage ring: stages x (A tile + BN x 128 B)
constexpr int TF_SMEM = TF_DATA_BYTES + 1024 /* barriers */ + 1024 /* alignment slack */;

struct TfConv {
  int ntaps; int toff[5]; int wtap[5];
  int kb0, kb1;                 // 32-channel blocks of source 0 / source 1 (concatenated input)
  int Tp, rbox, R, N, BN;
  int stages, stage_bytes;      // ring depth (<= TF_MAX_STAGES) and bytes per stage (A tile + B tile)
  int Tout, ostride, ooff;      // output tensor [R, Tout, N]; GEMM row (r, j) is written to slot j * ostride + ooff
  const float* bias; float* out; int accum;
};

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(TF_THREADS, 1) tf32_conv_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                                                                   const __grid_constant__ CUtensorMap tmW, const TfConv P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TF_DATA_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TF_MAX_STAGES + 2);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * TF_MAX_STAGES, bar_acc = bar_empty + 8 * TF_MAX_STAGES;
  const int TF_STAGES = P.stages;
  const uint32_t TF_STAGE_BYTES = (uint32_t)P.stage_bytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows_valid = P.rbox * P.Tp;
  const uint32_t a_bytes = (uint32_t)rows_valid * 128u, b_bytes = (uint32_t)P.BN * 128u;
  const int r0 = blockIdx.x * P.rbox, n0 = blockIdx.y * P.BN;
  const int kbs = P.kb0 + P.kb1, n_st = P.ntaps * kbs;
  const uint32_t tm_cols = P.BN <= 32 ? 32u : (P.BN <= 64 ? 64u : (P.BN <= 128 ? 128u : 256u));
  if (tid == 0) {
    for (int i = 0; i < TF_STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), tm_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0; uint32_t ph = 0;
    for (int tap = 0; tap < P.ntaps; ++tap) {
      for (int kb = 0; kb < kbs; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        if (elect_one()) {
          const uint32_t dst = smem_base + (uint32_t)s * TF_STAGE_BYTES;
          mbar_arrive_expect_tx(bar_full + 8 * s, a_bytes + b_bytes);
          if (kb < P.kb0) tma_load_3d(dst, &tmA0, kb * 32, P.toff[tap], r0, bar_full + 8 * s);
          else tma_load_3d(dst, &tmA1, (kb - P.kb0) * 32, P.toff[tap], r0, bar_full + 8 * s);
          tma_load_3d(dst + TF_A_BYTES, &tmW, kb * 32, n0, P.wtap[tap], bar_full + 8 * s);
        }
        __syncwarp();
        if (++s == TF_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int s = 0; uint32_t ph = 0;
    const uint32_t idesc = make_idesc_tf32(128, P.BN);
    for (int st = 0; st < n_st; ++st) {
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_base + (uint32_t)s * TF_STAGE_BYTES;
        const uint64_t ad = make_desc_sw128(a_addr, 1024), bd = make_desc_sw128(a_addr + TF_A_BYTES, 1024);
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_tf32(tmem_base, ad + 2 * j, bd + 2 * j, idesc, (st | j) ? 1u : 0u);
        umma_commit(bar_empty + 8 * s);
        if (st == n_st - 1) umma_commit(bar_acc);
      }
      __syncwarp();
      if (++s == TF_STAGES) { s = 0; ph ^= 1u; }
    }
  } else {
    // ===================== epilogue: lane = GEMM row of quadrant (warp & 3) =====================
    const int q = warp & 3, m = q * 32 + lane;
    const int r = r0 + m / P.Tp, t = m % P.Tp;
    const bool valid = m < rows_valid && r < P.R;
    float* op = P.out + ((size_t)r * P.Tout + t * P.ostride + P.ooff) * P.N + n0;
    // accumulate mode (the residual 1x1 convolution's data gradient is added to the block's): the previous values do not depend on
    // the MMAs -- up to 64 columns are fetched BEFORE the accumulator wait (read-modify-write behind the wait cost 27 us per launch)
    float4 prev[16];
    const bool pre = P.accum && P.BN <= 64;
    if (pre && valid) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i * 4 < P.BN) prev[i] = *reinterpret_cast<const float4*>(op + i * 4);
    }
    mbar_wait(bar_acc, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < P.BN; c0 += 64) {
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        const int c = c0 + cc;
        if (c < P.BN) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
          float4 pv[4];
          if (P.accum && !pre && valid) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pv[j] = *reinterpret_cast<const float4*>(op + c + 4 * j);
          }
          tmem_wait_ld();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              if (P.bias) {
                const float4 b = *reinterpret_cast<const float4*>(P.bias + n0 + c + j);
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              if (P.accum) {
                const float4 p = pre ? prev[(cc + j) >> 2] : pv[j >> 2];
                o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
              }
              *reinterpret_cast<float4*>(op + c + j) = o;
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tm_cols); }
}


// ------------------------------------------------------------------------------------------------
// weight gradient on the tensor pipe:  dW[tap][ci][co] = sum_{r,t} in[r, t + toff[tap], ci] * dOut[r, t, co]
// The reduction index (r, t) is the K dimension of the MMA, the channels are M (ci) and N (co): both operands are MN-MAJOR, i.e. the
// channels-last activation tensors are used as they lie -- a 3-D TMA box {32 channels, T' slots, rbox rows} is one 128B-swizzled
// "atom column" (32 MN elements x rbox*T' K rows, rows 128 B apart, 32-byte-atom swizzle repeating every 4 rows); the CTA tile of 128 ci x 64 co is 4 + 2
// such columns, 16 KB apart (the descriptor's leading-dimension offset).  One MMA (K = 8) consumes one 8-row group, so rbox * T' is a
// multiple of 8; slots outside [0, T') and rows beyond R are zero-filled by the TMA unit and add nothing.
// grid = (ci tiles, co tiles, taps x splits): every CTA reduces its share of the row boxes and writes a partial [ci][co] tile in the
// layout twreduce_kernel sums (fixed order: bit-reproducible).
// ------------------------------------------------------------------------------------------------
constexpr int TW_STAGES = 2;
constexpr int TW_COL = 16384;                          // one atom column: up to 128 K rows x 128 B
constexpr int TW_STAGE_BYTES = 6 * TW_COL;
constexpr int TW_SMEM = TW_STAGES * TW_STAGE_BYTES + 1024 + 1024;

struct TfWgrad {
  int ntaps; int toff[5];
  int c0, cin, cout;
  int Tp, rbox, R, nbox, bps, splits;
  int d_toff;                   // time coordinate of dOut's first slot (the output phase of a transposed convolution)
  float* part;
};

// MN-major operand of 32-bit elements: the "128-byte swizzle with 32-byte atoms" layout (descriptor layout type 1, TMA swizzle
// 128B_ATOM_32B; cute: Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> over 128 B x 4 K rows).  LBO = bytes between 32-element atom columns,
// SBO = bytes between 4-row (K) groups.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                               // SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(TF_THREADS, 1) tf32_wgrad_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                                                                    const __grid_constant__ CUtensorMap tmD, const TfWgrad P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TW_STAGES * TW_STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TW_STAGES + 2);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * TW_STAGES, bar_acc = bar_empty + 8 * TW_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = P.rbox * P.Tp;
  const int ci0 = blockIdx.x * 128, co0 = blockIdx.y * 64;
  const int tap = blockIdx.z % P.ntaps, split = blockIdx.z / P.ntaps;
  const int box_lo = split * P.bps, box_hi = min(P.nbox, box_lo + P.bps);
  const int a_cols = min(4, (P.cin - ci0) / 32);
  if (tid == 0) {
    for (int i = 0; i < TW_STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int bx = box_lo; bx < box_hi; ++bx) {
      mbar_wait(bar_empty + 8 * s, ph ^ 1u);
      if (elect_one()) {
        const uint32_t dst = smem_base + (uint32_t)s * TW_STAGE_BYTES;
        mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)(a_cols + 2) * (uint32_t)rows * 128u);
        for (int a = 0; a < a_cols; ++a) {
          const int c = ci0 + a * 32;
          if (c < P.c0) tma_load_3d(dst + a * TW_COL, &tmA0, c, P.toff[tap], bx * P.rbox, bar_full + 8 * s);
          else tma_load_3d(dst + a * TW_COL, &tmA1, c - P.c0, P.toff[tap], bx * P.rbox, bar_full + 8 * s);
        }
        tma_load_3d(dst + 4 * TW_COL, &tmD, co0, P.d_toff, bx * P.rbox, bar_full + 8 * s);
        tma_load_3d(dst + 5 * TW_COL, &tmD, co0 + 32, P.d_toff, bx * P.rbox, bar_full + 8 * s);
      }
      __syncwarp();
      if (++s == TW_STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    int s = 0; uint32_t ph = 0;
    const uint32_t idesc = make_idesc_tf32(128, 64) | (1u << 15) | (1u << 16);        // A and B MN-major
    const int groups = rows >> 3;
    bool first = true;
    for (int bx = box_lo; bx < box_hi; ++bx) {
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = smem_base + (uint32_t)s * TW_STAGE_BYTES;
        const uint64_t ad = make_desc_mn_sw128(a_addr, TW_COL, 512), bd = make_desc_mn_sw128(a_addr + 4 * TW_COL, TW_COL, 512);
        for (int g = 0; g < groups; ++g) {
          umma_tf32(tmem_base, ad + (uint64_t)(g * 64), bd + (uint64_t)(g * 64), idesc, first ? 0u : 1u);
          first = false;
        }
        umma_commit(bar_empty + 8 * s);
        if (bx == box_hi - 1) umma_commit(bar_acc);
      }
      first = false;
      __syncwarp();
      if (++s == TW_STAGES) { s = 0; ph ^= 1u; }
    }
  } else {
    const int q = warp & 3, ci = ci0 + q * 32 + lane;
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    float* op = P.part + (((size_t)split * P.ntaps + tap) * P.cin + ci) * P.cout + co0;
    for (int c = 0; c < 64; c += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tmem_wait_ld();
      if (ci < P.cin) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(op + c + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                               __uint_as_float(v[j + 3]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TfState {
  EncodeTiledFn enc = nullptr;
  bool attr_set = false, attr_set_w = false;
  std::map<std::tuple<const void*, int, int, int, int, int, int, int>, CUtensorMap> maps;     // (base, d0, d1, d2, box1, box2, swizzle, stride1)
};

static TfState* tf_of(CldHandle* h) {
  if (!h->train_tc) h->train_tc = new TfState();
  return reinterpret_cast<TfState*>(h->train_tc);
}
void train_tc_destroy(CldHandle* h) {
  delete reinterpret_cast<TfState*>(h->train_tc);
  h->train_tc = nullptr;
}

// 3-D fp32 tensor {d0 (contiguous), d1, d2}, box {32, b1, b2}, 128-byte swizzle, zero fill outside
// b1 = elements LOADED along dimension 1, taken every `es1`-th (traversal stride: the stride-2 convolutions)
static int tf_map(CldHandle* h, const float* base, int d0, int d1, int d2, int b1, int b2, const CUtensorMap** out, bool atom32 = false,
                  int es1 = 1) {
  TfState* st = tf_of(h);
  if (!st->enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return fail(h, CLD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    st->enc = (EncodeTiledFn)fn;
  }
  const auto key = std::make_tuple((const void*)base, d0, d1, d2, b1, b2, (int)atom32, es1);
  auto it = st->maps.find(key);
  if (it == st->maps.end()) {
    if (st->maps.size() > 4096) st->maps.clear();
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)d0 * 4, (cuuint64_t)d0 * d1 * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)(b1 * es1), (cuuint32_t)b2}, es[3] = {1, (cuuint32_t)es1, 1};
    const CUresult r = st->enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return fail(h, CLD_ERR_CUDA, "cuTensorMapEncodeTiled rejected a {%d,%d,%d} fp32 tensor with box {32,%d,%d}: %d", d0, d1, d2, b1, b2, (int)r);
    it = st->maps.emplace(key, m).first;
  }
  *out = &it->second;
  return 0;
}

bool tfconv_supported(int c0, int c1, int N, int Tp) {
  return c0 % 32 == 0 && c1 % 32 == 0 && c0 > 0 && N % 16 == 0 && N >= 16 && Tp >= 1 && Tp <= 128;
}

// out[r, j * ostride + ooff, 0:N) (+)= bias + sum_i  in[r, j * tstride + toff[i], :] @ W[wtap[i]]^T  for j in [0, Tp);
// in = [R, Ta, c0 (+ c1)], out = [R, Tout, N], W plane = [N][K] (K = c0 + c1 contiguous), `planes` planes
int tfconv_launch(CldHandle* h, const float* in0, int c0, const float* in1, int c1, int Ta, int tstride, int Tp, const float* w, int planes,
                  int ntaps, const int* wtap, const int* toff, const float* bias, float* out, int Tout, int ostride, int ooff, int N, int accum,
                  int R, cudaStream_t s) {
  TfState* st = tf_of(h);
  if (!st->attr_set) {
    CLD_CUDA_OK(h, cudaFuncSetAttribute(tf32_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM));
    st->attr_set = true;
  }
  TfConv P;
  P.ntaps = ntaps;
  for (int i = 0; i < 5; ++i) { P.toff[i] = i < ntaps ? toff[i] : 0; P.wtap[i] = i < ntaps ? wtap[i] : 0; }
  P.kb0 = c0 / 32; P.kb1 = c1 / 32; P.Tp = Tp; P.R = R; P.N = N;
  P.rbox = 128 / Tp;
  if (P.rbox > R) P.rbox = R;
  if (P.rbox > 256) P.rbox = 256;
  const int mtiles = (R + P.rbox - 1) / P.rbox;
  // N tile: the largest that still gives every SM a CTA (the A box is re-read once per N tile: large batches want wide tiles, the
  // 128-row minibatch wants many CTAs)
  P.BN = (N % 64 == 0 && mtiles * (N / 64) >= 120) ? 64 : (N % 32 == 0 ? 32 : 16);
  for (int bn = TF_MAX_BN; bn > 64; bn >>= 1)
    if (N % bn == 0 && mtiles * (N / bn) >= h->num_sms) { P.BN = bn; break; }
  P.stage_bytes = TF_A_BYTES + P.BN * 128;
  P.stages = TF_DATA_BYTES / P.stage_bytes;
  if (P.stages > TF_MAX_STAGES) P.stages = TF_MAX_STAGES;
  P.bias = bias; P.out = out; P.accum = accum; P.Tout = Tout; P.ostride = ostride; P.ooff = ooff;
  const CUtensorMap *mA0, *mA1, *mW;
  int rc;
  if ((rc = tf_map(h, in0, c0, Ta, R, Tp, P.rbox, &mA0, false, tstride))) return rc;
  mA1 = mA0;
  if (c1 > 0 && (rc = tf_map(h, in1, c1, Ta, R, Tp, P.rbox, &mA1, false, tstride))) return rc;
  if ((rc = tf_map(h, w, c0 + c1, N, planes, P.BN, 1, &mW))) return rc;
  dim3 grid(mtiles, N / P.BN);
  tf32_conv_kernel<<<grid, TF_THREADS, TF_SMEM, s>>>(*mA0, *mA1, *mW, P);
  CLD_LAUNCH_OK(h, "tf32_conv_kernel");
  return 0;
}

// rows of one box: the largest rbox with rbox * Tp <= 128 and rbox * Tp a multiple of 8 (one MMA consumes 8 reduction rows)
static int tfwgrad_rbox(int Tp) {
  for (int rb = 128 / Tp; rb >= 1; --rb)
    if ((rb * Tp) % 8 == 0) return rb;
  return 0;
}
bool tfwgrad_supported(int c0, int c1, int cout, int Tp, int R) {
  const int rb = Tp >= 1 && Tp <= 128 ? tfwgrad_rbox(Tp) : 0;
  return c0 > 0 && c0 % 32 == 0 && c1 % 32 == 0 && cout % 64 == 0 && rb > 0 && rb <= R;
}

// partials of dW[tap][ci][co] = sum_{r, j < Tp} in[r, j * a_stride + toff[tap], ci] * dout[r, j * d_stride + d_toff, co]
// (in = [R, Ta, c0 (+ c1)], dout = [R, Td, cout]) -> part [splits][ntaps][cin][cout]; *splits_out for the reduce
int tfwgrad_launch(CldHandle* h, const float* in0, int c0, const float* in1, int c1, int Ta, int a_stride, int Tp, const float* dout, int Td,
                   int d_stride, int d_toff, int cout, int ntaps, const int* toff, float* part, size_t part_floats, int max_splits, int R,
                   int* splits_out, cudaStream_t s) {
  TfState* st = tf_of(h);
  if (!st->attr_set_w) {
    CLD_CUDA_OK(h, cudaFuncSetAttribute(tf32_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM));
    st->attr_set_w = true;
  }
  TfWgrad P;
  P.ntaps = ntaps;
  for (int i = 0; i < 5; ++i) P.toff[i] = i < ntaps ? toff[i] : 0;
  P.c0 = c0; P.cin = c0 + c1; P.cout = cout; P.Tp = Tp; P.R = R; P.part = part; P.d_toff = d_toff;
  P.rbox = tfwgrad_rbox(Tp);
  P.nbox = (R + P.rbox - 1) / P.rbox;
  const int tiles = ((P.cin + 127) / 128) * (cout / 64) * ntaps;
  int splits = (2 * h->num_sms + tiles - 1) / tiles;
  if (splits > P.nbox) splits = P.nbox;
  if (splits > max_splits) splits = max_splits;
  const size_t wsz = (size_t)ntaps * P.cin * cout;
  if ((size_t)splits * wsz > part_floats) splits = (int)(part_floats / wsz);
  if (splits < 1) return fail(h, CLD_ERR_UNSUPPORTED, "weight-gradient scratch too small");
  P.bps = (P.nbox + splits - 1) / splits;
  splits = (P.nbox + P.bps - 1) / P.bps;
  P.splits = splits;
  const CUtensorMap *mA0, *mA1, *mD;
  int rc;
  if ((rc = tf_map(h, in0, c0, Ta, R, Tp, P.rbox, &mA0, true, a_stride))) return rc;
  mA1 = mA0;
  if (c1 > 0 && (rc = tf_map(h, in1, c1, Ta, R, Tp, P.rbox, &mA1, true, a_stride))) return rc;
  if ((rc = tf_map(h, dout, cout, Td, R, Tp, P.rbox, &mD, true, d_stride))) return rc;
  dim3 grid((P.cin + 127) / 128, cout / 64, ntaps * splits);
  tf32_wgrad_kernel<<<grid, TF_THREADS, TW_SMEM, s>>>(*mA0, *mA1, *mD, P);
  CLD_LAUNCH_OK(h, "tf32_wgrad_kernel");
  *splits_out = splits;
  return 0;
}

}  // namespace cld
