// Context encoder of the sampling path: ContextEncoder.forward (reference models/context_utils.py:40-61) on sm_100a.
//   map_encoder      : torchvision ResNet-18 with a 34-channel 7x7 stem, eval-mode BatchNorm, fc 512 -> 256
//                      (src/tbsim/models/base_models.py:573-607, diffuser_helpers.py:297-348)
//   agent_state_enc. : MLP 4 -> 64 -> 64 -> 64 with LayerNorm + ReLU (base_models.py:58-66)
//   process_cond_mlp : MLP 320 -> 320 -> 320 -> 256 -> 256 -> 256 with LayerNorm + ReLU
//
// Every convolution is ONE kernel, `conv_tma_kernel`: an implicit GEMM on the tcgen05 tensor pipe, fed by the TMA unit.
//   GEMM tile    = a BOX of 128 output pixels (bw x bh pixels of bn images); activations are NHWC bf16, so the 64 input
//                  channels of one filter tap are 128 contiguous bytes = one row of a 128B-swizzled K-major operand tile
//   k-block      = (filter tap, 64-channel panel) = ONE 4-D tensor-map box of the input, shifted by the tap; elements
//                  outside the image are zero-filled by the TMA unit (= the convolution's padding); no im2col buffer
//   weights      = pre-packed per (N tile, k-block) as swizzled [N][64] bf16 images, fetched by one bulk copy (UBLKCP)
//   accumulators = TMEM, two buffers of up to 256 columns: the epilogue of tile i (BatchNorm scale/shift, residual add,
//                  ReLU, bf16 NHWC store) overlaps the MMAs of tile i+1; persistent CTAs, one per SM
// The 7x7 stem (34 input channels) uses the same kernel: the raster is converted once to NHWC bf16 with a 36-channel
// pitch and 3 zero pixels left of every row, so for one filter row the 7 taps x 36 channels of an output pixel are 504
// CONTIGUOUS bytes; the TMA unit sees an OVERLAPPING view {256 elements, 112 output columns at a 144-byte stride, 224
// rows, B} and the four 64-element boxes of a view row are the k-blocks of that filter row (K = 7 x 256, not 49 x 64).
// Max-pool, the raster conversion and the MLP head are HBM-bound SIMT kernels.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/cld_b200.h"
#include "tc_common.cuh"

using namespace cld::tc;

namespace {

constexpr int CT_A_BYTES = 16384;      // 128 rows x 128 B
constexpr int CT_MAX_STAGES = 8;
constexpr int CT_TAIL = 4096 + 256 + 256;    // scale/bias [2][512] fp32 + barriers + TMEM slot + stage table
constexpr int IMG_C = 34, IMG_CP = 36, IMG_HW = 224, IMG_WP = 232, IMG_WOFF = 3, IMG_RUN = 256;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 32-byte global accesses (LDG/STG.256): a full sector per thread instead of two half-sector transactions
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ------------------------------------------------------------------------------------------------
// implicit-GEMM convolution + BatchNorm (+ residual) (+ ReLU); A operand moved by the TMA unit (cp.async.bulk.tensor, SASS UTMALDG).
// A GEMM tile is a BOX of output pixels (bw x bh pixels of bn images, bw*bh*bn = 128), so the A k-block of filter tap
// (dy, dx) and channel panel p is ONE 4-D tensor-map box of the NHWC input: {64 channels from p*64, bw pixels from
// ow0*s + dx - pad (traversal stride s), bh rows from oh0*s + dy - pad, bn images}; out-of-image elements are zero-filled by
// the TMA unit (= the convolution's padding) and the box lands in the 128B-swizzled K-major layout the MMA reads.
// Stem: the raster is stored [B,224,232,36] (3 zero pixels left, 5 right) and described to the TMA unit as an OVERLAPPING
// view {256 elements, 112 output columns at a 144-byte stride, 224 rows, B}: row (ow) of the view is the 7-tap x 36-channel
// run of output column ow, and its four 64-element boxes per filter row are the k-blocks.
// One elected thread issues both copies of a stage (A box + weight image) on one mbarrier.
// ------------------------------------------------------------------------------------------------
constexpr int CT2_THREADS = 192;       // warps 0-3 epilogue, 4 MMA issuer, 5 TMA producer
// A pipeline STAGE holds one activation box and the weight images of up to 4 filter taps that read it:
//   * plain stage (E = 1): the box is the 128-pixel tile shifted by one tap;
//   * grouped stage (3x3 stride-1 convolutions with tall tiles, and the stem): the box carries E-1 extra pixel rows
//     ("slices": bw pixels x bn images = slice_bytes), and filter row e of the group reads the SAME shared-memory box
//     through a descriptor that starts e slices further down -- the tensor map orders the box {channels, x, image, y} so
//     that a y step is one slice for every image of the tile.  L2 -> SM traffic per tile drops by E*bh/(bh+E-1).
//   * MT > 1 (stem): the CTA tile is MT m-tiles stacked in y (box rows MT*bh + E-1), every weight image feeds MT MMAs.
struct ConvT {
  __nv_bfloat16* out; const __nv_bfloat16* res;
  const uint8_t* wblob; const float* scale; const float* bias;
  int OH, OW, Cout, n_kb, NT, n_nt, n_mt, nb, relu, stages;
  int sx, sy, lbw, lbh, TX, TY;
  int e0, e1, e_split;                             // sub-blocks per stage: e0 for stages < e_split, e1 after (MMA issuer: no table read)
  int yb;                                          // box operand order: 1 = {run, x, image, y} (grouped stages), 0 = {run, x, y, image}
  int n_st, MT, a_bytes, slice_bytes, w_slots;     // stages per tile, m-tiles per CTA tile, A box bytes, weight images reserved per stage
  uint16_t stt[80];         // per stage: [0,4) channel box | [4,8) dx + 8 | [8,12) dy of sub-block 0 + 8 | [12,15) sub-blocks E
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

// keep a loop-invariant value in a register (the kernel parameters live in the constant bank and would otherwise be re-read
// inside the issue loop, behind the barrier wait)
__device__ __forceinline__ uint32_t pin(uint32_t x) { asm volatile("" : "+r"(x)); return x; }

template <int EMAX, int MT>
__global__ void __launch_bounds__(CT2_THREADS, 1) conv_tma_kernel(const __grid_constant__ CUtensorMap tmA, const ConvT P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int S = P.stages, NT = P.NT;
  const uint32_t b_bytes = (uint32_t)NT * 128u, a_bytes = (uint32_t)P.a_bytes, stage_bytes = a_bytes + (uint32_t)P.w_slots * b_bytes;
  uint8_t* tail = smem + (size_t)S * stage_bytes;
  float* sc_s = reinterpret_cast<float*>(tail);
  float* bi_s = sc_s + 512;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 4096);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 4096 + 8 * (2 * CT_MAX_STAGES + 4));
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * CT_MAX_STAGES;
  const uint32_t bar_accf = bar_empty + 8 * CT_MAX_STAGES, bar_acce = bar_accf + 16;
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint32_t* stt_s = reinterpret_cast<uint32_t*>(tail + 4096 + 256);      // stage table, two entries per word
  for (int i = tid; i < P.Cout; i += CT2_THREADS) { sc_s[i] = P.scale[i]; bi_s[i] = P.bias[i]; }
  if (tid < 40) stt_s[tid] = (uint32_t)P.stt[2 * tid] | ((uint32_t)P.stt[2 * tid + 1] << 16);
  if (tid == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_accf, 1); mbar_init(bar_accf + 8, 1);
    mbar_init(bar_acce, 4); mbar_init(bar_acce + 8, 4);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = P.n_mt * P.n_nt, n_st = P.n_st;
  const int bw = 1 << P.lbw, bh = 1 << P.lbh, bn = 128 >> (P.lbw + P.lbh);

  if (warp == 5) {
    // ===================== TMA producer: activation box + weight images per stage =====================
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int mt = tile / P.n_nt, nt = tile - mt * P.n_nt;
      const int tx = mt % P.TX, t2 = mt / P.TX, ty = t2 % P.TY, tb = t2 / P.TY;
      const int cx0 = tx * bw * P.sx, cy0 = ty * bh * MT * P.sy, cb0 = tb * bn;
      const uint8_t* wsrc = P.wblob + (size_t)nt * P.n_kb * b_bytes;
      for (int st = 0; st < n_st; ++st) {
        const uint32_t e = (stt_s[st >> 1] >> ((st & 1) * 16)) & 0xffffu;
        const uint32_t w_bytes = ((e >> 12) & 7u) * b_bytes;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        if (elect_one()) {
          const uint32_t dst = smem_base + (uint32_t)s * stage_bytes;
          mbar_arrive_expect_tx(bar_full + 8 * s, a_bytes + w_bytes);
          const int cx = cx0 + (int)((e >> 4) & 15u) - 8, cy = cy0 + (int)((e >> 8) & 15u) - 8;
          tma_load_4d(dst, &tmA, (int)(e & 15u) * 64, cx, P.yb ? cb0 : cy, P.yb ? cy : cb0, bar_full + 8 * s);
          bulk_g2s(dst + a_bytes, wsrc, w_bytes, bar_full + 8 * s);
        }
        __syncwarp();
        wsrc += w_bytes;
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    int s = 0; uint32_t ph = 0, ti = 0;
    const uint32_t idesc = pin(make_idesc_bf16(128, NT));
    const uint32_t slice16 = pin((uint32_t)P.slice_bytes >> 4), mt16 = pin(slice16 * (uint32_t)bh);
    const uint32_t stage16 = pin(stage_bytes >> 4), a16 = pin(a_bytes >> 4), b16 = pin(b_bytes >> 4), nt_cols = pin((uint32_t)NT);
    const int e0 = (int)pin((uint32_t)P.e0), e1 = (int)pin((uint32_t)P.e1);
    const int e_split = (int)pin((uint32_t)P.e_split), n_st_r = (int)pin((uint32_t)n_st), S_r = (int)pin((uint32_t)S);
    const uint64_t desc0 = make_desc_sw128(smem_base, 1024);       // + (byte offset >> 4) selects a tile inside the CTA's shared memory
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const uint32_t buf = ti & 1u, aph = (ti >> 1) & 1u;
      mbar_wait(bar_acce + 8 * buf, aph ^ 1u);
      tc_fence_after();
      const uint32_t d_addr = tmem_base + buf * 256u;
      for (int st = 0; st < n_st_r; ++st) {
        const uint64_t ad0 = desc0 + (uint64_t)((uint32_t)s * stage16), bd0 = ad0 + a16;
        const int E = EMAX == 1 ? 1 : (st < e_split ? e0 : e1);      // stem: 4 filter rows in the even group, 3 in the odd one
        mbar_wait(bar_full + 8 * s, ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int e = 0; e < EMAX; ++e) {
            if (e < E) {
              const uint64_t bd = bd0 + (uint64_t)((uint32_t)e * b16);
#pragma unroll
              for (int m = 0; m < MT; ++m) {
                const uint64_t ad = ad0 + (uint64_t)((uint32_t)e * slice16 + (uint32_t)m * mt16);
                const uint32_t d = d_addr + (uint32_t)m * nt_cols;
                umma_bf16(d, ad, bd, idesc, (st | e) != 0 ? 1u : 0u);
                umma_bf16(d, ad + 2, bd + 2, idesc, 1u);
                umma_bf16(d, ad + 4, bd + 4, idesc, 1u);
                umma_bf16(d, ad + 6, bd + 6, idesc, 1u);
              }
            }
          }
          umma_commit(bar_empty + 8 * s);
        }
        __syncwarp();
        if (++s == S_r) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(bar_accf + 8 * buf);
      __syncwarp();
    }
  } else {
    // ===================== epilogue: warp q owns TMEM lanes 32q .. 32q+31 =====================
    const int q = warp;
    uint32_t ti = 0;
    const int r = q * 32 + lane;
    // box order {x, image, y}: row r = (hi * bn + bi) * bw + wi;  {x, y, image}: r = (bi * bh + hi) * bw + wi
    const int wi = r & (bw - 1);
    const int bi = P.yb ? (r >> P.lbw) & (bn - 1) : r >> (P.lbw + P.lbh);
    const int hi = P.yb ? r >> (7 - P.lbh) : (r >> P.lbw) & (bh - 1);
    const bool has_res = P.res != nullptr;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int mt = tile / P.n_nt, nt = tile - mt * P.n_nt;
      const int tx = mt % P.TX, t2 = mt / P.TX, ty = t2 % P.TY, tb = t2 / P.TY;
      const uint32_t buf = ti & 1u, aph = (ti >> 1) & 1u;
      const int b = tb * bn + bi, ow = tx * bw + wi;
      const bool mv = b < P.nb;
      for (int m = 0; m < MT; ++m) {
        const int oh = (ty * MT + m) * bh + hi;
        const size_t orow = (((size_t)b * P.OH + oh) * P.OW + ow) * P.Cout + (size_t)nt * NT;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u + (uint32_t)(m * NT);
        for (int h0 = 0; h0 < NT; h0 += 128) {
          // the residual does not depend on the MMAs: fetch it (up to 128 channels of this row) before waiting for the accumulator
          uint4 rr[16];
          if (has_res && mv) {
#pragma unroll
            for (int j = 0; j < 16; j += 2)
              if (h0 + j * 8 < NT) ldg256(P.res + orow + h0 + j * 8, rr[j], rr[j + 1]);
          }
          if (h0 == 0 && m == 0) { mbar_wait(bar_accf + 8 * buf, aph); tc_fence_after(); }
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int c0 = h0 + cc * 32;
            if (c0 < NT) {
              uint32_t v[32];
              tmem_ld32(taddr + c0, v);
              tmem_wait_ld();
              if (mv) {
                const float4* sc4 = reinterpret_cast<const float4*>(sc_s + nt * NT + c0);
                const float4* bi4 = reinterpret_cast<const float4*>(bi_s + nt * NT + c0);
                uint4 pk[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  const float4 s0 = sc4[2 * g], s1 = sc4[2 * g + 1], o0 = bi4[2 * g], o1 = bi4[2 * g + 1];
                  float y[8];
                  y[0] = fmaf(__uint_as_float(v[g * 8 + 0]), s0.x, o0.x); y[1] = fmaf(__uint_as_float(v[g * 8 + 1]), s0.y, o0.y);
                  y[2] = fmaf(__uint_as_float(v[g * 8 + 2]), s0.z, o0.z); y[3] = fmaf(__uint_as_float(v[g * 8 + 3]), s0.w, o0.w);
                  y[4] = fmaf(__uint_as_float(v[g * 8 + 4]), s1.x, o1.x); y[5] = fmaf(__uint_as_float(v[g * 8 + 5]), s1.y, o1.y);
                  y[6] = fmaf(__uint_as_float(v[g * 8 + 6]), s1.z, o1.z); y[7] = fmaf(__uint_as_float(v[g * 8 + 7]), s1.w, o1.w);
                  if (has_res) {
                    const uint4 r4 = rr[cc * 4 + g];
                    y[0] += bf_lo(r4.x); y[1] += bf_hi(r4.x); y[2] += bf_lo(r4.y); y[3] += bf_hi(r4.y);
                    y[4] += bf_lo(r4.z); y[5] += bf_hi(r4.z); y[6] += bf_lo(r4.w); y[7] += bf_hi(r4.w);
                  }
                  if (P.relu) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
                  }
                  pk[g] = make_uint4(pack2(y[0], y[1]), pack2(y[2], y[3]), pack2(y[4], y[5]), pack2(y[6], y[7]));
                }
                stg256(P.out + orow + c0, pk[0], pk[1]);
                stg256(P.out + orow + c0 + 16, pk[2], pk[3]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// CTA-PAIR variant (cta_group::2) for the plain-stage layers with wide N (layers 3 and 4, the stride-2 convolutions): a
// cluster of two CTAs computes two adjacent m-tiles of the same N tile with ONE tcgen05.mma.cta_group::2 per K step (M = 256,
// 128 rows per CTA).  Each CTA stages its own activation box and only HALF of the weight image (N/2 rows); the instruction
// reads both halves, so the weight bytes per CTA and k-block halve (48 -> 32 KB at N = 256) -- the layers are bound by the
// L2 -> SM ingest.  Protocol: both producers' TMA copies complete on the LEADER's (rank 0) full barrier; the leader's MMA
// thread issues for the pair and commits with a multicast arrive onto the empty / accumulator-full barriers of both CTAs; the
// epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier (remote mbarrier arrive for rank 1).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t leader_bar) {
  asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(leader_bar) : "memory");
}
__global__ void __launch_bounds__(CT2_THREADS, 1) conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                                   const ConvT P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int S = P.stages, NT = P.NT;
  const uint32_t bh_bytes = (uint32_t)NT * 64u;                       // this CTA's half of a weight image: NT/2 rows x 128 B
  const uint32_t a_bytes = CT_A_BYTES, stage_bytes = a_bytes + bh_bytes;
  uint8_t* tail = smem + (size_t)S * stage_bytes;
  float* sc_s = reinterpret_cast<float*>(tail);
  float* bi_s = sc_s + 512;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 4096);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 4096 + 8 * (2 * CT_MAX_STAGES + 4));
  uint32_t* stt_s = reinterpret_cast<uint32_t*>(tail + 4096 + 256);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * CT_MAX_STAGES;
  const uint32_t bar_accf = bar_empty + 8 * CT_MAX_STAGES, bar_acce = bar_accf + 16;
  const uint32_t smem_base = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  for (int i = tid; i < P.Cout; i += CT2_THREADS) { sc_s[i] = P.scale[i]; bi_s[i] = P.bias[i]; }
  if (tid < 40) stt_s[tid] = (uint32_t)P.stt[2 * tid] | ((uint32_t)P.stt[2 * tid + 1] << 16);
  if (tid == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_accf, 1); mbar_init(bar_accf + 8, 1);
    mbar_init(bar_acce, 8); mbar_init(bar_acce + 8, 8);                // 4 epilogue warps of each CTA (used in the leader only)
    fence_barrier_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // both CTAs' barriers exist before any remote arrive / remote complete_tx
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_pair_tiles = ((P.n_mt + 1) >> 1) * P.n_nt, n_st = P.n_st;
  const int cid = (int)(blockIdx.x >> 1), n_clusters = (int)(gridDim.x >> 1);
  const int bw = 1 << P.lbw, bh = 1 << P.lbh, bn = 128 >> (P.lbw + P.lbh);

  if (warp == 5) {
    // ===================== TMA producer (both CTAs): own activation box + own half of the weight image =====================
    int s = 0; uint32_t ph = 0;
    const uint32_t full0 = mapa_shared(bar_full, 0);                     // the leader's full barriers
    for (int pt = cid; pt < n_pair_tiles; pt += n_clusters) {
      const int mp = pt / P.n_nt, nt = pt - mp * P.n_nt;
      const int mt = 2 * mp + (int)rank;
      const int tx = mt % P.TX, t2 = mt / P.TX, ty = t2 % P.TY, tb = t2 / P.TY;
      const int cx0 = tx * bw * P.sx, cy0 = ty * bh * P.sy, cb0 = tb * bn;
      const int wrow0 = nt * n_st * NT + (int)rank * (NT >> 1);          // first weight row of this CTA's half, k-block 0
      for (int st = 0; st < n_st; ++st) {
        const uint32_t e = (stt_s[st >> 1] >> ((st & 1) * 16)) & 0xffffu;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        if (elect_one()) {
          const uint32_t dst = smem_base + (uint32_t)s * stage_bytes;
          if (leader) mbar_expect_tx_only(bar_full + 8 * s, 2u * (a_bytes + bh_bytes));
          const int cx = cx0 + (int)((e >> 4) & 15u) - 8, cy = cy0 + (int)((e >> 8) & 15u) - 8;
          tma2_load_4d(dst, &tmA, (int)(e & 15u) * 64, cx, P.yb ? cb0 : cy, P.yb ? cy : cb0, full0 + 8 * s);
          tma2_load_2d(dst + a_bytes, &tmW, 0, wrow0 + st * NT, full0 + 8 * s);
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer (leader only): one instruction drives both CTAs =====================
    if (leader) {
      int s = 0; uint32_t ph = 0, ti = 0;
      const uint32_t idesc = pin(make_idesc_bf16(256, NT));
      const uint32_t stage16 = pin(stage_bytes >> 4), a16 = pin(a_bytes >> 4);
      const int n_st_r = (int)pin((uint32_t)n_st), S_r = (int)pin((uint32_t)S);
      const uint64_t desc0 = make_desc_sw128(smem_base, 1024);
      for (int pt = cid; pt < n_pair_tiles; pt += n_clusters, ++ti) {
        const uint32_t buf = ti & 1u, aph = (ti >> 1) & 1u;
        mbar_wait(bar_acce + 8 * buf, aph ^ 1u);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + buf * 256u;
        for (int st = 0; st < n_st_r; ++st) {
          const uint64_t ad = desc0 + (uint64_t)((uint32_t)s * stage16), bd = ad + a16;
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
            umma2_bf16(d_addr, ad, bd, idesc, st != 0 ? 1u : 0u);
            umma2_bf16(d_addr, ad + 2, bd + 2, idesc, 1u);
            umma2_bf16(d_addr, ad + 4, bd + 4, idesc, 1u);
            umma2_bf16(d_addr, ad + 6, bd + 6, idesc, 1u);
            umma2_commit_pair(bar_empty + 8 * s);
          }
          __syncwarp();
          if (++s == S_r) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) umma2_commit_pair(bar_accf + 8 * buf);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (both CTAs): own 128 rows =====================
    const int q = warp;
    uint32_t ti = 0;
    const int r = q * 32 + lane;
    const int wi = r & (bw - 1);
    const int bi = P.yb ? (r >> P.lbw) & (bn - 1) : r >> (P.lbw + P.lbh);
    const int hi = P.yb ? r >> (7 - P.lbh) : (r >> P.lbw) & (bh - 1);
    const bool has_res = P.res != nullptr;
    const uint32_t acce0 = mapa_shared(bar_acce, 0);
    for (int pt = cid; pt < n_pair_tiles; pt += n_clusters, ++ti) {
      const int mp = pt / P.n_nt, nt = pt - mp * P.n_nt;
      const int mt = 2 * mp + (int)rank;
      const int tx = mt % P.TX, t2 = mt / P.TX, ty = t2 % P.TY, tb = t2 / P.TY;
      const uint32_t buf = ti & 1u, aph = (ti >> 1) & 1u;
      const int b = tb * bn + bi, ow = tx * bw + wi, oh = ty * bh + hi;
      const bool mv = b < P.nb;
      const size_t orow = (((size_t)b * P.OH + oh) * P.OW + ow) * P.Cout + (size_t)nt * NT;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u;
      for (int h0 = 0; h0 < NT; h0 += 128) {
        uint4 rr[16];
        if (has_res && mv) {
#pragma unroll
          for (int j = 0; j < 16; j += 2)
            if (h0 + j * 8 < NT) ldg256(P.res + orow + h0 + j * 8, rr[j], rr[j + 1]);
        }
        if (h0 == 0) { mbar_wait(bar_accf + 8 * buf, aph); tc_fence_after(); }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c0 = h0 + cc * 32;
          if (c0 < NT) {
            uint32_t v[32];
            tmem_ld32(taddr + c0, v);
            tmem_wait_ld();
            if (mv) {
              const float4* sc4 = reinterpret_cast<const float4*>(sc_s + nt * NT + c0);
              const float4* bi4 = reinterpret_cast<const float4*>(bi_s + nt * NT + c0);
              uint4 pk[4];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float4 s0 = sc4[2 * g], s1 = sc4[2 * g + 1], o0 = bi4[2 * g], o1 = bi4[2 * g + 1];
                float y[8];
                y[0] = fmaf(__uint_as_float(v[g * 8 + 0]), s0.x, o0.x); y[1] = fmaf(__uint_as_float(v[g * 8 + 1]), s0.y, o0.y);
                y[2] = fmaf(__uint_as_float(v[g * 8 + 2]), s0.z, o0.z); y[3] = fmaf(__uint_as_float(v[g * 8 + 3]), s0.w, o0.w);
                y[4] = fmaf(__uint_as_float(v[g * 8 + 4]), s1.x, o1.x); y[5] = fmaf(__uint_as_float(v[g * 8 + 5]), s1.y, o1.y);
                y[6] = fmaf(__uint_as_float(v[g * 8 + 6]), s1.z, o1.z); y[7] = fmaf(__uint_as_float(v[g * 8 + 7]), s1.w, o1.w);
                if (has_res) {
                  const uint4 r4 = rr[cc * 4 + g];
                  y[0] += bf_lo(r4.x); y[1] += bf_hi(r4.x); y[2] += bf_lo(r4.y); y[3] += bf_hi(r4.y);
                  y[4] += bf_lo(r4.z); y[5] += bf_hi(r4.z); y[6] += bf_lo(r4.w); y[7] += bf_hi(r4.w);
                }
                if (P.relu) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
                }
                pk[g] = make_uint4(pack2(y[0], y[1]), pack2(y[2], y[3]), pack2(y[4], y[5]), pack2(y[6], y[7]));
              }
              stg256(P.out + orow + c0, pk[0], pk[1]);
              stg256(P.out + orow + c0 + 16, pk[2], pk[3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acce0 + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // nobody leaves (or frees tensor memory) while the peer may still touch this CTA
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------
// weight / parameter packing
// ------------------------------------------------------------------------------------------------
// w [Cout][Cin][KH][KW] fp32 -> per (N tile, k-block) swizzled [NT][64] bf16 images; k-block kb covers filter row
// tap.t[kb] & 15, filter column (>> 4) & 15 and channel box >> 8 (stem: the box is 64 elements of the 7-tap x 36-channel run)
struct KbTaps { uint16_t t[80]; };
__global__ void ctx_pack_conv_kernel(uint8_t* __restrict__ dst, const float* __restrict__ w, int Cout, int Cin, int KH, int KW,
                                     int NT, int n_kb, int stem, long long total, const KbTaps tap) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx & 63);
  long long t = idx >> 6;
  const int nr = (int)(t % NT); t /= NT;
  const int kb = (int)(t % n_kb);
  const int nt = (int)(t / n_kb);
  const int n = nt * NT + nr;
  const int dy = tap.t[kb] & 15, dx = (tap.t[kb] >> 4) & 15, box = tap.t[kb] >> 8;
  float v = 0.f;
  if (stem) {
    const int kk = box * 64 + k, dxp = kk / IMG_CP, ch = kk % IMG_CP;
    if (dxp < KW && ch < Cin) v = w[(((size_t)n * Cin + ch) * KH + dy) * KW + dxp];
  } else {
    const int ci = box * 64 + k;
    if (ci < Cin) v = w[(((size_t)n * Cin + ci) * KH + dy) * KW + dx];
  }
  *reinterpret_cast<__nv_bfloat16*>(dst + ((size_t)nt * n_kb + kb) * NT * 128 + sw128_off(nr, k >> 3) + (k & 7) * 2) = __float2bfloat16_rn(v);
}

// eval-mode BatchNorm as y = x * scale + shift
__global__ void ctx_fold_bn_kernel(float* scale, float* shift, const float* g, const float* b, const float* rm, const float* rv, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = g[i] / sqrtf(rv[i] + 1e-5f);
  scale[i] = s;
  shift[i] = b[i] - rm[i] * s;
}

__global__ void ctx_transpose_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int cols) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int r = idx / cols, c = idx - r * cols;
  dst[(size_t)c * rows + r] = src[idx];
}

// ------------------------------------------------------------------------------------------------
// HBM-bound helpers
// ------------------------------------------------------------------------------------------------
// image [B,34,224,224] fp32 (NCHW) -> [B,224,232,36] bf16 (NHWC; channels 34,35 zero; pixel w stored at column w + 3, the other
// columns zero); one block per image row, thread = pixel (17 independent coalesced loads of channel pairs in flight)
__global__ void __launch_bounds__(256) ctx_image_to_nhwc_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) uint32_t tile[IMG_WP * IMG_CP / 2];
  const int b = blockIdx.x / IMG_HW, h = blockIdx.x % IMG_HW, w = threadIdx.x;
  const float* src = img + ((size_t)b * IMG_C * IMG_HW + h) * IMG_HW + w;
  constexpr size_t CS = (size_t)IMG_HW * IMG_HW;
  for (int i = threadIdx.x; i < (IMG_WP - IMG_HW) * IMG_CP / 2; i += 256) {       // left / right padding columns
    const int col = i / (IMG_CP / 2), k = i % (IMG_CP / 2);
    tile[(col < IMG_WOFF ? col : col + IMG_HW) * (IMG_CP / 2) + k] = 0u;
  }
  if (w < IMG_HW) {
    float v[IMG_C];
#pragma unroll
    for (int c = 0; c < IMG_C; ++c) v[c] = __ldg(src + c * CS);
    uint32_t* dst = tile + (w + IMG_WOFF) * (IMG_CP / 2);
#pragma unroll
    for (int c = 0; c < IMG_C / 2; ++c) dst[c] = pack2(v[2 * c], v[2 * c + 1]);
    dst[IMG_C / 2] = 0u;
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * IMG_HW + h) * IMG_WP * IMG_CP);
  const uint4* ts = reinterpret_cast<const uint4*>(tile);
  for (int i = threadIdx.x; i < IMG_WP * IMG_CP / 8; i += 256) dst[i] = ts[i];
}

// History rasterisation fused with the raster layout (rasterize_agents, reference src/tbsim/utils/trajdata_utils.py:123-156):
// pass 1 writes the map layers into channels 31..33 and zeros everywhere else, one block per image row
__global__ void __launch_bounds__(256) ctx_raster_base_kernel(const float* __restrict__ maps, __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) uint32_t tile[IMG_WP * IMG_CP / 2];
  const int b = blockIdx.x / IMG_HW, h = blockIdx.x % IMG_HW, w = threadIdx.x;
  for (int i = threadIdx.x; i < IMG_WP * IMG_CP / 2; i += 256) tile[i] = 0u;
  __syncthreads();
  if (w < IMG_HW) {
    constexpr size_t CS = (size_t)IMG_HW * IMG_HW;
    const float* src = maps + ((size_t)b * 3 * IMG_HW + h) * IMG_HW + w;
    const float m0 = __ldg(src), m1 = __ldg(src + CS), m2 = __ldg(src + 2 * CS);
    uint32_t* dst = tile + (w + IMG_WOFF) * (IMG_CP / 2);
    dst[15] = pack2(0.f, m0);            // channels 30 (last history frame, scattered later), 31
    dst[16] = pack2(m1, m2);             // channels 32, 33
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * IMG_HW + h) * IMG_WP * IMG_CP);
  const uint4* ts = reinterpret_cast<const uint4*>(tile);
  for (int i = threadIdx.x; i < IMG_WP * IMG_CP / 8; i += 256) dst[i] = ts[i];
}
// pass 2 (others, value -1) and pass 3 (ego = agent 0, value +1, launched after so that it wins): one thread per
// (image, agent, history frame); pixel = round(clip(raster_from_agent * p)); the first and the last pixel of the raster are
// never written (trajdata_utils.py:148-149: they collect the unavailable / out-of-range positions)
__global__ void ctx_raster_scatter_kernel(const float* __restrict__ pos, const uint8_t* __restrict__ mask, const float* __restrict__ rfa,
                                          __nv_bfloat16* __restrict__ out, int nb, int A, int T, int ego) {
  const int na = ego ? 1 : A - 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nb * na * T) return;
  const int t = idx % T, a = (idx / T) % na + (ego ? 0 : 1), b = idx / (T * na);
  const size_t pi = ((size_t)b * A + a) * T + t;
  float x = 0.f, y = 0.f;
  if (mask[pi]) {
    const float px = pos[pi * 2], py = pos[pi * 2 + 1];
    const float* M = rfa + (size_t)b * 9;
    x = __fadd_rn(__fadd_rn(__fmul_rn(px, M[0]), __fmul_rn(py, M[1])), M[2]);
    y = __fadd_rn(__fadd_rn(__fmul_rn(px, M[3]), __fmul_rn(py, M[4])), M[5]);
  }
  const int ix = (int)rintf(fminf(fmaxf(x, 0.f), (float)(IMG_HW - 1))), iy = (int)rintf(fminf(fmaxf(y, 0.f), (float)(IMG_HW - 1)));
  const int flat = iy * IMG_HW + ix;
  if (flat == 0 || flat == IMG_HW * IMG_HW - 1) return;
  out[(((size_t)b * IMG_HW + iy) * IMG_WP + ix + IMG_WOFF) * IMG_CP + t] = __float2bfloat16_rn(ego ? 1.f : -1.f);
}
// the workspace raster back as the reference's image [B,34,224,224] fp32
__global__ void ctx_raster_export_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int nb) {
  const long long total = (long long)nb * IMG_C * IMG_HW * IMG_HW;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = (int)(idx % IMG_HW);
  long long t = idx / IMG_HW;
  const int h = (int)(t % IMG_HW); t /= IMG_HW;
  const int c = (int)(t % IMG_C);
  const int b = (int)(t / IMG_C);
  out[idx] = __bfloat162float(in[(((size_t)b * IMG_HW + h) * IMG_WP + w + IMG_WOFF) * IMG_CP + c]);
}

__device__ __forceinline__ uint32_t bmax2(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
// MaxPool2d(3, stride 2, padding 1) on NHWC bf16, 16 channels (one 32-byte sector) per thread
__global__ void ctx_maxpool_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int H, int W, int C,
                                   int OH, int OW) {
  const long long total = (long long)B * OH * OW * (C >> 4);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cc = (int)(idx % (C >> 4));
  long long t = idx / (C >> 4);
  const int ow = (int)(t % OW); t /= OW;
  const int oh = (int)(t % OH);
  const int b = (int)(t / OH);
  uint4 m0 = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u), m1 = m0;   // -inf pairs
  for (int dy = 0; dy < 3; ++dy) {
    const int ih = oh * 2 - 1 + dy;
    if (ih < 0 || ih >= H) continue;
    for (int dx = 0; dx < 3; ++dx) {
      const int iw = ow * 2 - 1 + dx;
      if (iw < 0 || iw >= W) continue;
      uint4 v0, v1;
      ldg256(in + (((size_t)b * H + ih) * W + iw) * C + cc * 16, v0, v1);
      m0.x = bmax2(m0.x, v0.x); m0.y = bmax2(m0.y, v0.y); m0.z = bmax2(m0.z, v0.z); m0.w = bmax2(m0.w, v0.w);
      m1.x = bmax2(m1.x, v1.x); m1.y = bmax2(m1.y, v1.y); m1.z = bmax2(m1.z, v1.z); m1.w = bmax2(m1.w, v1.w);
    }
  }
  stg256(out + (((size_t)b * OH + oh) * OW + ow) * C + cc * 16, m0, m1);
}

// debug / verification tap: NHWC bf16 -> NCHW fp32
__global__ void ctx_tap_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int B, int H, int W, int C) {
  const long long total = (long long)B * H * W * C;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  long long t = idx / C;
  const int w = (int)(t % W); t /= W;
  const int h = (int)(t % H);
  const int b = (int)(t / H);
  out[(((size_t)b * C + c) * H + h) * W + w] = __bfloat162float(in[idx]);
}

// ------------------------------------------------------------------------------------------------
// head: avg-pool + fc + agent-state MLP + process_cond MLP, 8 agents per CTA, fp32
// ------------------------------------------------------------------------------------------------
constexpr int HD_THREADS = 320, HD_LD = 512;       // AG agents per CTA (template): 8 amortises the weight reads of large batches, 1 fills the SMs for small ones
struct HeadP {
  const __nv_bfloat16* feat;      // [B,7,7,512] layer4 output
  const float* curr;              // [B,4]
  float* cond;                    // [B,256]
  float* map_feat;                // optional [B,256]
  int B, npix;
  // transposed weights [K][N], biases, LayerNorm scale / shift
  const float *s0w, *s0b, *s0g, *s0e, *s1w, *s1b, *s1g, *s1e, *s2w, *s2b;
  const float *fcw, *fcb;
  const float *p0w, *p0b, *p0g, *p0e, *p1w, *p1b, *p1g, *p1e, *p2w, *p2b, *p2g, *p2e, *p3w, *p3b, *p3g, *p3e, *p4w, *p4b;
};

// out[a][n] = sum_k in[a][k] * Wt[k][n] + b[n]   (thread = n, 8 agents share every weight read; K is a multiple of 4 and the
// activations are read as broadcast float4 -- two shared-memory instructions per k instead of eight)
template <int HD_AG>
__device__ __forceinline__ void hd_linear(const float* in, int ldi, const float* __restrict__ Wt, const float* __restrict__ b, float* out,
                                          int ldo, int K, int N) {
  const int n = threadIdx.x;
  if (n < N) {
    float acc[HD_AG];
#pragma unroll
    for (int a = 0; a < HD_AG; ++a) acc[a] = 0.f;
#pragma unroll 8
    for (int k = 0; k < K; k += 4) {
      const float w0 = __ldg(Wt + (size_t)k * N + n), w1 = __ldg(Wt + (size_t)(k + 1) * N + n);
      const float w2 = __ldg(Wt + (size_t)(k + 2) * N + n), w3 = __ldg(Wt + (size_t)(k + 3) * N + n);
#pragma unroll
      for (int a = 0; a < HD_AG; ++a) {
        const float4 x = *reinterpret_cast<const float4*>(in + a * ldi + k);
        acc[a] = fmaf(x.x, w0, fmaf(x.y, w1, fmaf(x.z, w2, fmaf(x.w, w3, acc[a]))));
      }
    }
    const float bb = b[n];
#pragma unroll
    for (int a = 0; a < HD_AG; ++a) out[a * ldo + n] = acc[a] + bb;
  }
  __syncthreads();
}
// LayerNorm (eps 1e-5, biased variance) + ReLU in place; warp a handles agent a
template <int HD_AG>
__device__ __forceinline__ void hd_ln_relu(float* x, int ld, const float* __restrict__ g, const float* __restrict__ e, int N) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < HD_AG) {
    float* row = x + warp * ld;
    float s = 0.f;
    for (int i = lane; i < N; i += 32) s += row[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / N;
    float v = 0.f;
    for (int i = lane; i < N; i += 32) { const float d = row[i] - mean; v = fmaf(d, d, v); }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / N + 1e-5f);
    for (int i = lane; i < N; i += 32) row[i] = fmaxf((row[i] - mean) * rstd * g[i] + e[i], 0.f);
  }
  __syncthreads();
}

template <int HD_AG>
__global__ void __launch_bounds__(HD_THREADS) ctx_head_kernel(const HeadP P) {
  __shared__ __align__(16) float bufA[HD_AG * HD_LD], bufB[HD_AG * HD_LD];
  const int a0 = blockIdx.x * HD_AG, tid = threadIdx.x;
  // average pool: bufA[a][c], c < 512
  for (int i = tid; i < HD_AG * 64; i += HD_THREADS) {
    const int a = i >> 6, cc = i & 63, ag = min(a0 + a, P.B - 1);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < P.npix; ++p) {
      const uint4 v = *reinterpret_cast<const uint4*>(P.feat + ((size_t)ag * P.npix + p) * 512 + cc * 8);
      acc[0] += bf_lo(v.x); acc[1] += bf_hi(v.x); acc[2] += bf_lo(v.y); acc[3] += bf_hi(v.y);
      acc[4] += bf_lo(v.z); acc[5] += bf_hi(v.z); acc[6] += bf_lo(v.w); acc[7] += bf_hi(v.w);
    }
    const float inv = 1.f / P.npix;
    for (int j = 0; j < 8; ++j) bufA[a * HD_LD + cc * 8 + j] = acc[j] * inv;
  }
  __syncthreads();
  // fc 512 -> 256 into bufB[a][64 + n]  (the concat layout: [state 0..63 | map 64..319])
  hd_linear<HD_AG>(bufA, HD_LD, P.fcw, P.fcb, bufB + 64, HD_LD, 512, 256);
  if (P.map_feat != nullptr) {
    for (int i = tid; i < HD_AG * 256; i += HD_THREADS) {
      const int a = i >> 8, n = i & 255;
      if (a0 + a < P.B) P.map_feat[(size_t)(a0 + a) * 256 + n] = bufB[a * HD_LD + 64 + n];
    }
  }
  // agent-state MLP: curr -> bufA
  if (tid < HD_AG * 4) { const int a = tid >> 2, ag = min(a0 + a, P.B - 1); bufA[a * HD_LD + (tid & 3)] = P.curr[(size_t)ag * 4 + (tid & 3)]; }
  __syncthreads();
  float* t1 = bufA + 64;          // scratch columns inside bufA (row stride HD_LD)
  hd_linear<HD_AG>(bufA, HD_LD, P.s0w, P.s0b, t1, HD_LD, 4, 64);
  hd_ln_relu<HD_AG>(t1, HD_LD, P.s0g, P.s0e, 64);
  float* t2 = bufA + 128;
  hd_linear<HD_AG>(t1, HD_LD, P.s1w, P.s1b, t2, HD_LD, 64, 64);
  hd_ln_relu<HD_AG>(t2, HD_LD, P.s1g, P.s1e, 64);
  hd_linear<HD_AG>(t2, HD_LD, P.s2w, P.s2b, bufB, HD_LD, 64, 64);            // state feature -> bufB[a][0..63]
  // process_cond_mlp on bufB[a][0..319]
  hd_linear<HD_AG>(bufB, HD_LD, P.p0w, P.p0b, bufA, HD_LD, 320, 320);
  hd_ln_relu<HD_AG>(bufA, HD_LD, P.p0g, P.p0e, 320);
  hd_linear<HD_AG>(bufA, HD_LD, P.p1w, P.p1b, bufB, HD_LD, 320, 320);
  hd_ln_relu<HD_AG>(bufB, HD_LD, P.p1g, P.p1e, 320);
  hd_linear<HD_AG>(bufB, HD_LD, P.p2w, P.p2b, bufA, HD_LD, 320, 256);
  hd_ln_relu<HD_AG>(bufA, HD_LD, P.p2g, P.p2e, 256);
  hd_linear<HD_AG>(bufA, HD_LD, P.p3w, P.p3b, bufB, HD_LD, 256, 256);
  hd_ln_relu<HD_AG>(bufB, HD_LD, P.p3g, P.p3e, 256);
  hd_linear<HD_AG>(bufB, HD_LD, P.p4w, P.p4b, bufA, HD_LD, 256, 256);
  for (int i = tid; i < HD_AG * 256; i += HD_THREADS) {
    const int a = i >> 8, n = i & 255;
    if (a0 + a < P.B) P.cond[(size_t)(a0 + a) * 256 + n] = bufA[a * HD_LD + n];
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct ConvLayer {
  int Cin_real, Cin, Cout, KH, KW, stride, pad, stem, panels, n_kb, NT, n_nt, stages;
  uint8_t* wblob = nullptr; float* scale = nullptr; float* shift = nullptr;
  // execution plan (fixed buffers): input / output / residual, input size, ReLU, TMA description of the input
  const __nv_bfloat16* in = nullptr; __nv_bfloat16* out = nullptr; const __nv_bfloat16* res = nullptr;
  int H = 0, relu = 0, lbw = 0, lbh = 0;
  int yb = 0, group = 0, MT = 1, n_st = 0, a_bytes = CT_A_BYTES, slice_bytes = 0, w_slots = 1;   // stage grouping (see ConvT)
  uint16_t stt[80];
  KbTaps taps;
  CUtensorMap tmap;
  int pair = 0;             // CTA-pair kernel (cta_group::2): needs tmap_w, the weight blob as a 2-D tensor {64, rows}
  CUtensorMap tmap_w;
};

std::string g_ctx_create_err;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct CldContext {
  int device = 0, num_sms = 0, max_agents = 0, chunk = 0;
  std::string err;
  std::vector<void*> allocs;
  ConvLayer conv[20];           // 0 stem; then per block conv1, conv2[, downsample] (state-dict order)
  int order[20];                // execution order: conv1, downsample, conv2
  const __nv_bfloat16* tap_buf[5] = {nullptr};
  float* head_w[30] = {nullptr};
  __nv_bfloat16 *img16 = nullptr, *stem_out = nullptr, *bufX = nullptr, *bufY = nullptr, *bufZ = nullptr, *bufD = nullptr;
  bool loaded = false;
  unsigned long long launches = 0;
  void* enc = nullptr;                   // cuTensorMapEncodeTiled
  double conv_flops_per_agent = 0.0;     // 2*MAC of the 20 convolutions as executed (padded K included)
};

namespace {

int cfail(CldContext* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_ctx_create_err = buf;
  return code;
}
#define CTX_CUDA_OK(c, expr)                                                                                     \
  do {                                                                                                           \
    cudaError_t _e = (expr);                                                                                     \
    if (_e != cudaSuccess) return cfail(c, CLD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define CTX_LAUNCH_OK(c, name)                                                                                   \
  do {                                                                                                           \
    cudaError_t _e = cudaGetLastError();                                                                         \
    if (_e != cudaSuccess) return cfail(c, CLD_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    ++(c)->launches;                                                                                             \
  } while (0)

template <typename T>
int calloc_dev(CldContext* c, T** p, size_t n) {
  void* q = nullptr;
  CTX_CUDA_OK(c, cudaMalloc(&q, (n ? n : 1) * sizeof(T)));
  c->allocs.push_back(q);
  *p = (T*)q;
  return 0;
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

void plan_conv(ConvLayer& L, int cin_real, int cout, int k, int stride, int pad, int stem) {
  L.Cin_real = cin_real; L.Cout = cout; L.KH = k; L.KW = k; L.stride = stride; L.pad = pad; L.stem = stem;
  L.Cin = stem ? IMG_CP : cin_real;
  L.panels = stem ? IMG_RUN / 64 : cin_real / 64;
  L.n_kb = stem ? k * L.panels : k * k * L.panels;
  const int nt_max = env_int("CLD_CTX_NT", 256);
  L.NT = cout < nt_max ? cout : nt_max;
  L.n_nt = cout / L.NT;
  L.stages = 0;
}

// stage plan + tensor map of a layer's input for conv_tma_kernel; returns false when the driver refuses the tensor map
bool plan_stages(EncodeTiledFn enc, ConvLayer& L, int chunk, bool group) {
  const int OH = (L.H + 2 * L.pad - L.KH) / L.stride + 1;
  switch (OH) {
    case 112: L.lbw = 4; L.lbh = 3; break;
    case 56: L.lbw = 3; L.lbh = 3; break;
    case 28: L.lbw = 2; L.lbh = 2; break;
    case 14: L.lbw = 1; L.lbh = 1; break;
    case 7: L.lbw = 0; L.lbh = 0; break;
    default: return false;
  }
  const cuuint32_t bw = 1u << L.lbw, bh = 1u << L.lbh, bn = 128u >> (L.lbw + L.lbh);
  L.slice_bytes = (int)(bw * bn * 128u);
  // ---- stages and the weight-image order that goes with them
  int emax = 1;
  L.group = 0; L.MT = 1; L.n_st = 0;
  int kb = 0;
  auto stage = [&](int cbox, int dxo, int dy0, int E) { L.stt[L.n_st++] = (uint16_t)(cbox | ((dxo + 8) << 4) | ((dy0 + 8) << 8) | (E << 12)); };
  if (L.stem && group) {
    // filter rows of equal parity read the same box: (even rows 0,2,4,6 | odd rows 1,3,5) x 4 channel-run boxes
    L.group = 1; L.MT = 2; emax = 4;
    for (int par = 0; par < 2; ++par)
      for (int kp = 0; kp < L.panels; ++kp) {
        const int E = par == 0 ? 4 : 3;
        stage(kp, 0, par - L.pad, E);
        for (int e = 0; e < E; ++e) L.taps.t[kb++] = (uint16_t)((2 * e + par) | (kp << 8));
      }
  } else if (!L.stem && group && L.KH == 3 && L.stride == 1 && bh >= 4 && (bh >= 8 || env_int("CLD_CTX_GROUP", 7) & 4)) {
    L.group = 1; emax = 3;
    for (int dx = 0; dx < 3; ++dx)
      for (int pn = 0; pn < L.panels; ++pn) {
        stage(pn, dx - L.pad, -L.pad, 3);
        for (int e = 0; e < 3; ++e) L.taps.t[kb++] = (uint16_t)(e | (dx << 4) | (pn << 8));
      }
  } else if (L.stem) {
    for (int dy = 0; dy < L.KH; ++dy)
      for (int kp = 0; kp < L.panels; ++kp) { stage(kp, 0, dy - L.pad, 1); L.taps.t[kb++] = (uint16_t)(dy | (kp << 8)); }
  } else {
    for (int dy = 0; dy < L.KH; ++dy)
      for (int dx = 0; dx < L.KW; ++dx)
        for (int pn = 0; pn < L.panels; ++pn) { stage(pn, dx - L.pad, dy - L.pad, 1); L.taps.t[kb++] = (uint16_t)(dy | (dx << 4) | (pn << 8)); }
  }
  if (kb != L.n_kb || L.n_st > 80) return false;
  const int rows_y = L.MT * (int)bh + emax - 1;            // box extent in y, in output rows
  L.a_bytes = rows_y * L.slice_bytes;
  L.w_slots = emax;
  const int stage_bytes = L.a_bytes + L.w_slots * L.NT * 128;
  int st = (216 * 1024) / stage_bytes;
  L.stages = st > CT_MAX_STAGES ? CT_MAX_STAGES : st;
  if (L.stages < 2) return false;
  // ---- tensor map; grouped stages need the y extent of the box outermost ({channel run, x, image, y}), plain stages use the
  // natural order {channel run, x, y, image}
  L.yb = (L.group && bn > 1) || env_int("CLD_CTX_YB", 0);
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], es[4];
  const int iy = L.yb ? 3 : 2, ib = L.yb ? 2 : 3;
  if (L.stem) {
    const cuuint64_t row = (cuuint64_t)IMG_WP * IMG_CP * 2;
    dims[0] = IMG_RUN; dims[1] = 112; dims[ib] = (cuuint64_t)chunk; dims[iy] = IMG_HW;
    strides[0] = 2 * IMG_CP * 2; strides[ib - 1] = row * IMG_HW; strides[iy - 1] = row;
    box[0] = 64; box[1] = bw; box[ib] = bn; box[iy] = (cuuint32_t)rows_y * 2;
    es[0] = 1; es[1] = 1; es[ib] = 1; es[iy] = 2;
  } else {
    const cuuint64_t px = (cuuint64_t)L.Cin * 2;
    dims[0] = (cuuint64_t)L.Cin; dims[1] = (cuuint64_t)L.H; dims[ib] = (cuuint64_t)chunk; dims[iy] = (cuuint64_t)L.H;
    strides[0] = px; strides[ib - 1] = px * L.H * L.H; strides[iy - 1] = px * L.H;
    box[0] = 64; box[1] = bw * L.stride; box[ib] = bn; box[iy] = (cuuint32_t)rows_y * L.stride;
    es[0] = 1; es[1] = (cuuint32_t)L.stride; es[ib] = 1; es[iy] = (cuuint32_t)L.stride;
  }
  const CUresult r = enc(&L.tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)L.in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int launch_conv(CldContext* c, const ConvLayer& L, int B, cudaStream_t s) {
  const int OH = (L.H + 2 * L.pad - L.KH) / L.stride + 1, OW = OH;
  const size_t smem = (size_t)L.stages * (L.a_bytes + L.w_slots * L.NT * 128) + CT_TAIL;
  ConvT P;
  P.out = L.out; P.res = L.res; P.wblob = L.wblob; P.scale = L.scale; P.bias = L.shift;
  P.OH = OH; P.OW = OW; P.Cout = L.Cout; P.n_kb = L.n_kb; P.NT = L.NT; P.n_nt = L.n_nt; P.nb = B; P.relu = L.relu; P.stages = L.stages;
  P.lbw = L.lbw; P.lbh = L.lbh; P.TX = OW >> L.lbw; P.TY = (OH >> L.lbh) / L.MT;
  const int bn = 128 >> (L.lbw + L.lbh);
  P.n_mt = P.TX * P.TY * ((B + bn - 1) / bn);
  P.sx = L.stem ? 1 : L.stride; P.sy = L.stride;
  P.e0 = (L.stt[0] >> 12) & 7; P.e1 = (L.stt[L.n_st - 1] >> 12) & 7; P.e_split = L.n_st / 2;
  P.yb = L.yb; P.n_st = L.n_st; P.MT = L.MT; P.a_bytes = L.a_bytes; P.slice_bytes = L.slice_bytes; P.w_slots = L.w_slots;
  memcpy(P.stt, L.stt, sizeof(P.stt));
  const int tiles = P.n_mt * P.n_nt;
  const int grid = tiles < c->num_sms ? tiles : c->num_sms;
  if (L.pair) {
    const int pair_tiles = ((P.n_mt + 1) / 2) * P.n_nt;
    int clusters = c->num_sms / 2;
    if (pair_tiles < clusters) clusters = pair_tiles;
    const int stage_b = CT_A_BYTES + L.NT * 64;
    int st = (216 * 1024) / stage_b;
    P.stages = st > CT_MAX_STAGES ? CT_MAX_STAGES : st;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters)); cfg.blockDim = dim3(CT2_THREADS);
    cfg.dynamicSmemBytes = (size_t)P.stages * stage_b + CT_TAIL; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_pair_kernel, L.tmap, L.tmap_w, P);
    if (e != cudaSuccess) return cfail(c, CLD_ERR_CUDA, "launch of conv_pair_kernel failed: %s", cudaGetErrorString(e));
    ++c->launches;
    return 0;
  }
  if (L.w_slots == 1 && L.MT == 1) conv_tma_kernel<1, 1><<<grid, CT2_THREADS, smem, s>>>(L.tmap, P);
  else if (L.w_slots == 3 && L.MT == 1) conv_tma_kernel<3, 1><<<grid, CT2_THREADS, smem, s>>>(L.tmap, P);
  else if (L.w_slots == 4 && L.MT == 2) conv_tma_kernel<4, 2><<<grid, CT2_THREADS, smem, s>>>(L.tmap, P);
  else return cfail(c, CLD_ERR_STATE, "launch_conv: no kernel variant for %d weight slots x %d m-tiles", L.w_slots, L.MT);
  CTX_LAUNCH_OK(c, "conv_tma_kernel");
  return 0;
}

}  // namespace

extern "C" {

int cld_context_create(int max_agents, CldContext** out) {
  if (!out || max_agents <= 0) return cfail(nullptr, CLD_ERR_ARG, "cld_context_create: bad arguments");
  *out = nullptr;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return cfail(nullptr, CLD_ERR_CUDA, "cld_context_create: no CUDA device");
  if (prop.major != 10) return cfail(nullptr, CLD_ERR_ARCH, "cld_context_create: device is sm_%d%d, this library is sm_100a only (no fallback)", prop.major, prop.minor);
  CldContext* c = new CldContext();
  c->device = dev; c->num_sms = prop.multiProcessorCount; c->max_agents = max_agents;
  const int chunk_max = env_int("CLD_CTX_CHUNK", 2048);
  c->chunk = max_agents < chunk_max ? max_agents : chunk_max;
  EncodeTiledFn enc = nullptr;
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      enc = (EncodeTiledFn)fn;
    if (!enc) { delete c; return cfail(nullptr, CLD_ERR_CUDA, "cld_context_create: cuTensorMapEncodeTiled is not available from the driver"); }
  }
  c->enc = (void*)enc;
  // plan: stem, then (conv1, conv2[, downsample]) per BasicBlock
  int li = 0;
  plan_conv(c->conv[li++], IMG_C, 64, 7, 2, 3, 1);
  int cin = 64;
  for (int l = 0; l < 4; ++l) {
    const int cout = 64 << l;
    for (int b = 0; b < 2; ++b) {
      const int stride = (l > 0 && b == 0) ? 2 : 1;
      plan_conv(c->conv[li++], cin, cout, 3, stride, 1, 0);
      plan_conv(c->conv[li++], cout, cout, 3, 1, 1, 0);
      if (stride == 2) plan_conv(c->conv[li++], cin, cout, 1, 2, 0, 0);
      cin = cout;
    }
  }
  const size_t n = (size_t)c->chunk;
  int rc = 0;
  // workspace: raster (NHWC bf16), stem output, one standing block buffer; the other three alias the raster region, which
  // is dead once the stem has run
  if ((rc = calloc_dev(c, &c->img16, n * IMG_HW * IMG_WP * IMG_CP)) || (rc = calloc_dev(c, &c->stem_out, n * 112 * 112 * 64)) ||
      (rc = calloc_dev(c, &c->bufX, n * 56 * 56 * 64))) {
    g_ctx_create_err = c->err;
    for (void* q : c->allocs) cudaFree(q);
    delete c;
    return rc;
  }
  c->bufY = c->img16;
  c->bufZ = c->img16 + n * 56 * 56 * 64;
  c->bufD = c->img16 + 2 * n * 56 * 56 * 64;
  // execution plan
  {
    int k = 0;
    ConvLayer& st = c->conv[0];
    st.in = c->img16; st.out = c->stem_out; st.res = nullptr; st.H = IMG_HW; st.relu = 1;
    c->order[k++] = 0;
    c->tap_buf[0] = c->bufX;
    __nv_bfloat16 *X = c->bufX, *Y = c->bufY, *Z = c->bufZ, *D = c->bufD;
    int l1 = 1, h = 56;
    for (int l = 0; l < 4; ++l) {
      for (int b = 0; b < 2; ++b) {
        const bool ds = l > 0 && b == 0;
        const int ho = ds ? h / 2 : h;
        ConvLayer &c1 = c->conv[l1], &c2 = c->conv[l1 + 1];
        c1.in = X; c1.out = Y; c1.res = nullptr; c1.H = h; c1.relu = 1;
        c->order[k++] = l1;
        c2.res = X;
        if (ds) {
          ConvLayer& cd = c->conv[l1 + 2];
          cd.in = X; cd.out = D; cd.res = nullptr; cd.H = h; cd.relu = 0;
          c->order[k++] = l1 + 2;
          c2.res = D;
        }
        c2.in = Y; c2.out = Z; c2.H = ho; c2.relu = 1;
        c->order[k++] = l1 + 1;
        l1 += ds ? 3 : 2;
        __nv_bfloat16* t = X; X = Z; Z = t;
        h = ho;
      }
      c->tap_buf[l + 1] = X;
    }
  }
  // CLD_CTX_GROUP (debug): bit 0 groups the stem's filter rows, bit 1 those of the stride-1 3x3 convolutions with 8-row tiles
  // (layer1), bit 2 also those with 4-row tiles (layer2)
  const int group_mask = env_int("CLD_CTX_GROUP", 7);
  for (int i = 0; i < 20; ++i) {
    ConvLayer& L = c->conv[i];
    const bool want = L.stem ? (group_mask & 1) : (group_mask & 2);
    if (!plan_stages(enc, L, c->chunk, want) && !(want && plan_stages(enc, L, c->chunk, false))) {
      for (void* q : c->allocs) cudaFree(q);
      delete c;
      return cfail(nullptr, CLD_ERR_CUDA, "cld_context_create: cuTensorMapEncodeTiled rejected the activation tensor of convolution %d", i);
    }
    if (env_int("CLD_CTX_VERBOSE", 0))
      fprintf(stderr, "[cld_context] conv %2d: Cin %3d Cout %3d k%d s%d in %3d | NT %3d stages %d x %5.1f KB, %2d stages/tile, grouped %d, MT %d\n", i,
              L.Cin_real, L.Cout, L.KH, L.stride, L.H, L.NT, L.stages, (L.a_bytes + L.w_slots * L.NT * 128) / 1024.0, L.n_st, L.group, L.MT);
  }
  cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  cudaFuncSetAttribute(conv_tma_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  cudaFuncSetAttribute(conv_tma_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  cudaFuncSetAttribute(conv_tma_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  *out = c;
  return 0;
}

void cld_context_destroy(CldContext* c) {
  if (!c) return;
  for (void* p : c->allocs) cudaFree(p);
  delete c;
}

const char* cld_context_last_error(const CldContext* c) { return c ? c->err.c_str() : g_ctx_create_err.c_str(); }

unsigned long long cld_context_launch_count(const CldContext* c) { return c ? c->launches : 0ull; }

double cld_context_conv_flops(const CldContext* c) { return c ? c->conv_flops_per_agent : 0.0; }

/* 130 fp32 device tensors: ContextEncoder.state_dict() order without the `num_batches_tracked` entries. */
int cld_context_load(CldContext* c, const float* const* p, const int64_t* numels, int n, void* stream) {
  if (!c || !p) return CLD_ERR_ARG;
  if (n != 130) return cfail(c, CLD_ERR_ARG, "cld_context_load: expected 130 tensors, got %d", n);
  cudaStream_t s = (cudaStream_t)stream;
  int idx = 10, rc = 0;
  auto want = [&](int i, long long ne) -> int {
    if (numels && numels[i] != ne) return cfail(c, CLD_ERR_ARG, "cld_context_load: tensor %d has %lld elements, expected %lld", i, (long long)numels[i], ne);
    return 0;
  };
  double flops = 0.0;
  for (int li = 0; li < 20; ++li) {
    ConvLayer& L = c->conv[li];
    // state-dict order inside a BasicBlock is conv1, bn1, conv2, bn2, downsample; execution order is the same
    const float* w = p[idx];
    if ((rc = want(idx, (long long)L.Cout * L.Cin_real * L.KH * L.KW))) return rc;
    const float *g = p[idx + 1], *b = p[idx + 2], *rm = p[idx + 3], *rv = p[idx + 4];
    for (int j = 1; j <= 4; ++j) if ((rc = want(idx + j, L.Cout))) return rc;
    idx += 5;
    const size_t bytes = (size_t)L.n_nt * L.n_kb * L.NT * 128;
    if (!L.wblob) {
      if ((rc = calloc_dev(c, &L.wblob, bytes))) return rc;
      if ((rc = calloc_dev(c, &L.scale, L.Cout))) return rc;
      if ((rc = calloc_dev(c, &L.shift, L.Cout))) return rc;
    }
    // CTA-pair kernel: the long-K plain-stage layers with N = 256 tiles (the 3x3 convolutions of layers 3 and 4; measured -4..-8 %
    // per launch, while the short 1x1 / stride-2 launches lose to the pair protocol).  CLD_CTX_PAIR=0 disables it, =2 uses it for
    // every plain-stage layer with N tiles >= 128.  The weight blob is described as a 2-D tensor whose boxes are the half images
    // (NT/2 rows) each CTA of a pair stages.
    L.pair = 0;
    const int pair_mode = env_int("CLD_CTX_PAIR", 1);
    if (!L.stem && !L.group && ((pair_mode == 1 && L.NT == 256 && L.n_kb >= 36) || (pair_mode >= 2 && L.NT >= 128))) {
      cuuint64_t dims[2] = {64, (cuuint64_t)L.n_nt * L.n_kb * L.NT};
      cuuint64_t strides[1] = {128};
      cuuint32_t box[2] = {64, (cuuint32_t)(L.NT / 2)}, es[2] = {1, 1};
      const CUresult r = ((EncodeTiledFn)c->enc)(&L.tmap_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)L.wblob, dims, strides, box, es,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      L.pair = r == CUDA_SUCCESS ? 1 : 0;
    }
    const long long total = (long long)L.n_nt * L.n_kb * L.NT * 64;
    ctx_pack_conv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(L.wblob, w, L.Cout, L.Cin_real, L.KH, L.KW, L.NT, L.n_kb, L.stem, total,
                                                                        L.taps);
    CTX_LAUNCH_OK(c, "ctx_pack_conv_kernel");
    ctx_fold_bn_kernel<<<(L.Cout + 127) / 128, 128, 0, s>>>(L.scale, L.shift, g, b, rm, rv, L.Cout);
    CTX_LAUNCH_OK(c, "ctx_fold_bn_kernel");
  }
  // conv FLOPs per agent as executed (2 * M * N * K with the padded K)
  {
    int h = IMG_HW;
    int li = 0;
    auto add = [&](const ConvLayer& L, int hin) { const int oh = (hin + 2 * L.pad - L.KH) / L.stride + 1; flops += 2.0 * oh * oh * L.Cout * (double)L.n_kb * 64; return oh; };
    h = add(c->conv[li++], h);      // stem -> 112
    h = 56;                         // max-pool
    for (int l = 0; l < 4; ++l)
      for (int b = 0; b < 2; ++b) {
        const bool ds = l > 0 && b == 0;
        const int ho = add(c->conv[li++], h);
        add(c->conv[li++], ho);
        if (ds) add(c->conv[li++], h);
        h = ho;
      }
    c->conv_flops_per_agent = flops;
  }
  // head weights: transposed copies [K][N] of the Linear layers, vectors as they are
  // tensor indices: agent_state_encoder 0..9, fc idx..idx+1, process_cond_mlp after
  const int fc = idx, pm = idx + 2;
  if (pm + 18 != 130) return cfail(c, CLD_ERR_STATE, "cld_context_load: internal index mismatch (%d)", pm);
  // (source index, out features, in features) for matrices; vectors copied verbatim
  const int mats[][3] = {{0, 64, 4}, {4, 64, 64}, {8, 64, 64}, {fc, 256, 512}, {pm + 0, 320, 320}, {pm + 4, 320, 320}, {pm + 8, 256, 320},
                         {pm + 12, 256, 256}, {pm + 16, 256, 256}};
  const int mat_slot[] = {0, 4, 8, 10, 12, 16, 20, 24, 28};
  for (int i = 0; i < 9; ++i) {
    const int src = mats[i][0], rows = mats[i][1], cols = mats[i][2];
    if ((rc = want(src, (long long)rows * cols))) return rc;
    float*& d = c->head_w[mat_slot[i]];
    if (!d && (rc = calloc_dev(c, &d, (size_t)rows * cols))) return rc;
    ctx_transpose_kernel<<<(rows * cols + 255) / 256, 256, 0, s>>>(d, p[src], rows, cols);
    CTX_LAUNCH_OK(c, "ctx_transpose_kernel");
  }
  // vectors: slot <- source
  const int vecs[][3] = {{1, 1, 64}, {2, 2, 64}, {3, 3, 64}, {5, 5, 64}, {6, 6, 64}, {7, 7, 64}, {9, 9, 64}, {11, fc + 1, 256},
                         {13, pm + 1, 320}, {14, pm + 2, 320}, {15, pm + 3, 320}, {17, pm + 5, 320}, {18, pm + 6, 320}, {19, pm + 7, 320},
                         {21, pm + 9, 256}, {22, pm + 10, 256}, {23, pm + 11, 256}, {25, pm + 13, 256}, {26, pm + 14, 256}, {27, pm + 15, 256},
                         {29, pm + 17, 256}};
  for (auto& v : vecs) {
    if ((rc = want(v[1], v[2]))) return rc;
    float*& d = c->head_w[v[0]];
    if (!d && (rc = calloc_dev(c, &d, (size_t)v[2]))) return rc;
    CTX_CUDA_OK(c, cudaMemcpyAsync(d, p[v[1]], sizeof(float) * v[2], cudaMemcpyDeviceToDevice, s));
  }
  c->loaded = true;
  return 0;
}

}  // extern "C"

namespace {
struct RasterSrc {           // exactly one of the two sources is set
  const float* image = nullptr;                                      // [B,34,224,224] fp32
  const float* maps = nullptr; const float* hist_pos = nullptr; const uint8_t* hist_mask = nullptr; const float* rfa = nullptr;
  int A = 0;                                                          // agents per history (ego first)
  float* image_out = nullptr;                                         // optional export of the rasterised image
};

int forward_impl(CldContext* c, const RasterSrc& src, const float* curr_states, int B, float* cond_feat, float* map_feat_out, int tap_stage,
                 float* tap_out, cudaStream_t s) {
  int rc = 0;
  constexpr int T = IMG_C - 3;
  for (int b0 = 0; b0 < B; b0 += c->chunk) {
    const int nb = (B - b0) < c->chunk ? (B - b0) : c->chunk;
    if (src.image) {
      ctx_image_to_nhwc_kernel<<<nb * IMG_HW, 256, 0, s>>>(src.image + (size_t)b0 * IMG_C * IMG_HW * IMG_HW, c->img16);
      CTX_LAUNCH_OK(c, "ctx_image_to_nhwc_kernel");
    } else {
      ctx_raster_base_kernel<<<nb * IMG_HW, 256, 0, s>>>(src.maps + (size_t)b0 * 3 * IMG_HW * IMG_HW, c->img16);
      CTX_LAUNCH_OK(c, "ctx_raster_base_kernel");
      const float* pos = src.hist_pos + (size_t)b0 * src.A * T * 2;
      const uint8_t* msk = src.hist_mask + (size_t)b0 * src.A * T;
      const float* rfa = src.rfa + (size_t)b0 * 9;
      if (src.A > 1) {
        const int n = nb * (src.A - 1) * T;
        ctx_raster_scatter_kernel<<<(n + 255) / 256, 256, 0, s>>>(pos, msk, rfa, c->img16, nb, src.A, T, 0);
        CTX_LAUNCH_OK(c, "ctx_raster_scatter_kernel");
      }
      ctx_raster_scatter_kernel<<<(nb * T + 255) / 256, 256, 0, s>>>(pos, msk, rfa, c->img16, nb, src.A, T, 1);
      CTX_LAUNCH_OK(c, "ctx_raster_scatter_kernel");
      if (src.image_out) {
        const long long total = (long long)nb * IMG_C * IMG_HW * IMG_HW;
        ctx_raster_export_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(c->img16, src.image_out + (size_t)b0 * IMG_C * IMG_HW * IMG_HW, nb);
        CTX_LAUNCH_OK(c, "ctx_raster_export_kernel");
      }
      if (!cond_feat) continue;            // rasterisation only
    }
    auto tap = [&](int stage, int h, int ch) -> int {
      if (tap_stage != stage || !tap_out) return 0;
      const long long total = (long long)nb * h * h * ch;
      ctx_tap_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(c->tap_buf[stage], tap_out, nb, h, h, ch);
      CTX_LAUNCH_OK(c, "ctx_tap_kernel");
      return 0;
    };
    for (int k = 0; k < 20; ++k) {
      const ConvLayer& L = c->conv[c->order[k]];
      if ((rc = launch_conv(c, L, nb, s))) return rc;
      if (k == 0) {
        const long long total = (long long)nb * 56 * 56 * 4;
        ctx_maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(c->stem_out, c->bufX, nb, 112, 112, 64, 56, 56);
        CTX_LAUNCH_OK(c, "ctx_maxpool_kernel");
        if ((rc = tap(0, 56, 64))) return rc;
      }
      // a layer ends after its 4th (layer1) / 5th (layer2..4) convolution
      if (k == 4 && (rc = tap(1, 56, 64))) return rc;
      if (k == 9 && (rc = tap(2, 28, 128))) return rc;
      if (k == 14 && (rc = tap(3, 14, 256))) return rc;
      if (k == 19 && (rc = tap(4, 7, 512))) return rc;
    }
    const __nv_bfloat16* X = c->tap_buf[4];
    HeadP hp;
    hp.feat = X; hp.curr = curr_states + (size_t)b0 * 4; hp.cond = cond_feat + (size_t)b0 * 256;
    hp.map_feat = map_feat_out ? map_feat_out + (size_t)b0 * 256 : nullptr;
    hp.B = nb; hp.npix = 49;
    float** w = c->head_w;
    hp.s0w = w[0]; hp.s0b = w[1]; hp.s0g = w[2]; hp.s0e = w[3]; hp.s1w = w[4]; hp.s1b = w[5]; hp.s1g = w[6]; hp.s1e = w[7]; hp.s2w = w[8]; hp.s2b = w[9];
    hp.fcw = w[10]; hp.fcb = w[11];
    hp.p0w = w[12]; hp.p0b = w[13]; hp.p0g = w[14]; hp.p0e = w[15]; hp.p1w = w[16]; hp.p1b = w[17]; hp.p1g = w[18]; hp.p1e = w[19];
    hp.p2w = w[20]; hp.p2b = w[21]; hp.p2g = w[22]; hp.p2e = w[23]; hp.p3w = w[24]; hp.p3b = w[25]; hp.p3g = w[26]; hp.p3e = w[27];
    hp.p4w = w[28]; hp.p4b = w[29];
    // agents per CTA: enough CTAs to fill the SMs for small batches, weight reads amortised over 8 agents for large ones
    const int ag = nb <= c->num_sms ? 1 : (nb <= 2 * c->num_sms ? 2 : (nb <= 8 * c->num_sms ? 4 : 8));
    const int hgrid = (nb + ag - 1) / ag;
    if (ag == 1) ctx_head_kernel<1><<<hgrid, HD_THREADS, 0, s>>>(hp);
    else if (ag == 2) ctx_head_kernel<2><<<hgrid, HD_THREADS, 0, s>>>(hp);
    else if (ag == 4) ctx_head_kernel<4><<<hgrid, HD_THREADS, 0, s>>>(hp);
    else ctx_head_kernel<8><<<hgrid, HD_THREADS, 0, s>>>(hp);
    CTX_LAUNCH_OK(c, "ctx_head_kernel");
  }
  return 0;
}
}  // namespace

extern "C" {

/* cond_feat = ContextEncoder.forward(data_batch)['cond_feat']  (models/context_utils.py:40-61).
 * image [B,34,224,224] fp32, curr_states [B,4] fp32 (x, y, vel, yaw: batch_utils.get_current_states) -> cond_feat [B,256].
 * map_feat_out (optional) [B,256] = the ResNet's fc output.  tap_stage >= 0 (debug, B <= chunk): copies the activation after
 * stage 0 (stem + max-pool), 1..4 (layer1..layer4) to tap_out as fp32 NCHW. */
int cld_context_forward(CldContext* c, const float* image, const float* curr_states, int B, float* cond_feat, float* map_feat_out,
                        int tap_stage, float* tap_out, void* stream) {
  if (!c || !image || !curr_states || !cond_feat || B <= 0) return c ? cfail(c, CLD_ERR_ARG, "cld_context_forward: bad arguments") : CLD_ERR_ARG;
  if (!c->loaded) return cfail(c, CLD_ERR_STATE, "cld_context_forward: weights not loaded");
  if (tap_stage >= 0 && B > c->chunk) return cfail(c, CLD_ERR_ARG, "cld_context_forward: taps need B <= %d", c->chunk);
  RasterSrc src;
  src.image = image;
  return forward_impl(c, src, curr_states, B, cond_feat, map_feat_out, tap_stage, tap_out, (cudaStream_t)stream);
}

/* The same from the un-rasterised inputs: rasterize_agents (src/tbsim/utils/trajdata_utils.py:123-156, called by
 * parse_node_centric :395-420) is fused with the raster layout, so the 6.8 MB/agent fp32 image never exists.
 *   maps [B,3,224,224] fp32 map layers; hist_pos [B,A,31,2] fp32 history positions in the ego frame, agent 0 = ego;
 *   hist_mask [B,A,31] bytes (availability); raster_from_agent [B,3,3] fp32.
 * cond_feat == NULL: rasterise only.  image_out (optional) [B,34,224,224] fp32 receives the image rasterize_agents returns. */
int cld_context_forward_history(CldContext* c, const float* maps, const float* hist_pos, const uint8_t* hist_mask, const float* raster_from_agent,
                                int num_hist_agents, const float* curr_states, int B, float* cond_feat, float* map_feat_out, float* image_out,
                                void* stream) {
  if (!c || !maps || !hist_pos || !hist_mask || !raster_from_agent || B <= 0 || num_hist_agents < 1)
    return c ? cfail(c, CLD_ERR_ARG, "cld_context_forward_history: bad arguments") : CLD_ERR_ARG;
  if (cond_feat && !curr_states) return cfail(c, CLD_ERR_ARG, "cld_context_forward_history: curr_states missing");
  if (cond_feat && !c->loaded) return cfail(c, CLD_ERR_STATE, "cld_context_forward_history: weights not loaded");
  RasterSrc src;
  src.maps = maps; src.hist_pos = hist_pos; src.hist_mask = hist_mask; src.rfa = raster_from_agent; src.A = num_hist_agents; src.image_out = image_out;
  return forward_impl(c, src, curr_states, B, cond_feat, map_feat_out, -1, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
