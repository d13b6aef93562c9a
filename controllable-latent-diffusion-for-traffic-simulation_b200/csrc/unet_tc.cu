// bf16 tensor-core denoiser: TemporalMapUnet.forward (reference src/tbsim/models/temporal.py:122-180) as ONE
// persistent sm_100a kernel.  A CTA owns a group of 8 batch rows and carries them through all 35
// convolutions with the activations resident in shared memory (bf16, channels-last, 128B-swizzled) and the
// accumulators in tensor memory; the 8.7 MB bf16 weight set is streamed from L2 through an 8 x 8 KB ring by the
// bulk-copy (TMA) unit.  Layout trick: GEMM row m = time_slot * 8 + batch_row, so a convolution tap is a shift
// by whole 1024-byte swizzle atoms of the SAME shared-memory tile (5 taps = 5 UMMA descriptors, no im2col),
// stride-2 convolutions are the same descriptor with SBO = 2048, the transposed convolution is two output
// phases.  GroupNorm statistics are CTA-local because a CTA holds whole rows; GroupNorm + Mish + time/cond bias
// + residual are fused into the TMEM -> register -> shared-memory epilogue that writes the next layer's A operand.
//
// Warp roles (320 threads): warp 0 weight producer, warp 1 MMA issuer (one elected thread), warps 2..9 epilogue.
#include <vector>

#include "tc_common.cuh"
#include "unet_tc.cuh"

namespace cld {
using namespace tc;

constexpr int TC_G = 8;
constexpr int TC_UNIT = 8192;
constexpr int TC_UNITS = 8;
constexpr int TC_THREADS = 320;
constexpr int TC_ARENA = 128 * 1024;
constexpr int TB_LD = 260;            // padded row stride of the time-bias tile (bank-conflict free float4 reads)
constexpr int TC_MAX_OPS = 32;
constexpr int TC_MAX_KBS = 640;

enum { EPI_GN_TB = 0, EPI_GN_RES_ACC = 1, EPI_GN_RES_ID = 2, EPI_BIAS = 3, EPI_UP = 4, EPI_GN = 5, EPI_OUT = 6 };

struct TcOp {
  int n, n_tiles;
  int tile_slot0[4], tile_lo[4], tile_hi[4];
  int sbo, slot_stride;
  int kb_first, n_kb, units, kb_bytes, w_first;
  int epi, cout, cpg, t_out, n_vt;
  int dst_off, dst_pitch;
  int par_off, tb_off, res_col;
  int zero_pitch, zero_npanels, zero_offB;     // zero_pitch == 0: no halo zeroing
  int save_skip;                               // -1 or byte offset inside the CTA's skip buffer
  int load_skip, load_off, load_pitch, load_npanels, load_T;   // load_skip: -1 or byte offset in the skip buffer
  int dbg_stage;
};
struct TcKb { int a_base, shift, w_off, acc_col, nk16, first; };   // host-side record; the device gets it packed in 32 bits
constexpr int TC_SMEM = TC_ARENA + TC_UNITS * TC_UNIT + 4096 + TC_G * TB_LD * 4 + 8192 + 512 + TC_MAX_OPS * (int)sizeof(TcOp) +
                        TC_MAX_KBS * 4 + 256 + 1024;

struct TcParams {
  const TcOp* ops; int n_ops; const uint32_t* kbs; int n_kbs;
  const uint8_t* wblob; const float* par; const float* tbias; const float* tvec; int tb_stride;
  const float* x; float* eps; int R, T, n_groups;
  uint8_t* skipbuf; int skip_stride;
  int zero0_pitch, zero0_npanels, zero0_offB;
  int dbg_stage; float* dbg_out;
};

struct TcState {
  std::vector<TcOp> ops;
  std::vector<TcKb> kbs;
  TcOp* d_ops = nullptr; uint32_t* d_kbs = nullptr;
  uint8_t* wblob = nullptr; size_t wblob_bytes = 0;
  float* par = nullptr; size_t par_floats = 0;
  uint8_t* skipbuf = nullptr; int skip_stride = 0; int grid = 0;
  // weight pointers recorded by tc_pack_* until tc_finalize
  struct Blk { int cin, cout; const float *c0w, *c0b, *g0, *b0, *c1w, *c1b, *g1, *b1, *rw, *rb; } blk[12];
  struct Rs { int ch; const float *w, *b; } down[2], up[2];
  const float *fw, *fb, *fg, *fbt, *f1w, *f1b;
  bool ready = false;
};

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mish_fast(float x) {
  // x * tanh(softplus(x)) = x * n / (n + 2) = x * (1 - 2 / (n + 2)),  n = e^x (e^x + 2); inf-safe (1/inf = 0)
  float e = __expf(x);
  float d = fmaf(e, e + 2.f, 2.f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return x * fmaf(r, -2.f, 1.f);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ void zero_halos(uint8_t* arena, int offB, int pitch, int npanels, int etid) {
  // per region: leading 2 slots of every panel + 2 tail slots after the last panel = (npanels+1) blocks of 2 KB
  const int blocks = 2 * (npanels + 1);
  for (int i = etid; i < blocks * 128; i += 256) {
    int blk = i >> 7, w = i & 127;
    int reg = blk / (npanels + 1), p = blk % (npanels + 1);
    uint4* dst = reinterpret_cast<uint4*>(arena + (reg ? offB : 0) + p * pitch) + w;
    *dst = make_uint4(0u, 0u, 0u, 0u);
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) unet_tc_kernel(const TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* arena = smem_raw;   // kept as a __shared__-space pointer so that ptxas emits LDS/STS, not generic LD/ST
  if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();   // the swizzle atoms need 1024-byte alignment
  uint8_t* ring = arena + TC_ARENA;
  float* par_s = reinterpret_cast<float*>(ring + TC_UNITS * TC_UNIT);   // [4][256] bias, gamma, beta, res bias
  float* tb_s = par_s + 1024;                                           // [8][TB_LD] time/cond bias per row
  float* st_s = tb_s + TC_G * TB_LD;                                    // [4 quadrants][8][32][2] partial sums
  float* mr_s = st_s + 2048;                                            // [8][8][2] mean, rstd per (row, group)
  TcOp* ops_s = reinterpret_cast<TcOp*>(mr_s + 128);                    // [TC_MAX_OPS]
  uint32_t* kbs_s = reinterpret_cast<uint32_t*>(ops_s + TC_MAX_OPS);    // [TC_MAX_KBS] packed k-block records
  uint64_t* bars = reinterpret_cast<uint64_t*>(kbs_s + TC_MAX_KBS);     // full[8], empty[8], act_ready, acc_ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 8);
  const uint32_t bar_act = smem_u32(bars + 16), bar_acc = smem_u32(bars + 17);

  for (int i = tid; i < P.n_ops * (int)(sizeof(TcOp) / 4); i += TC_THREADS)
    reinterpret_cast<uint32_t*>(ops_s)[i] = reinterpret_cast<const uint32_t*>(P.ops)[i];
  for (int i = tid; i < P.n_kbs; i += TC_THREADS) kbs_s[i] = P.kbs[i];
  if (tid == 0) {
    for (int i = 0; i < TC_UNITS; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_act, 1);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      uint32_t par_empty = 0;
      int pos = 0;
      for (int g = blockIdx.x; g < P.n_groups; g += gridDim.x) {
        for (int oi = 0; oi < P.n_ops; ++oi) {
          const TcOp* o = ops_s + oi;
          const int u = o->units, nkb = o->n_kb, bytes = o->kb_bytes;
          const uint8_t* src = P.wblob + (size_t)(unsigned)o->w_first;
          for (int k = 0; k < nkb; ++k) {
            pos = (pos + u - 1) & ~(u - 1);
            if (pos + u > TC_UNITS) pos = 0;
            for (int uu = pos; uu < pos + u; ++uu) {
              mbar_wait(bar_empty + 8 * uu, ((par_empty >> uu) & 1u) ^ 1u);
              par_empty ^= 1u << uu;
            }
            mbar_arrive_expect_tx(bar_full + 8 * pos, bytes);
            bulk_g2s(smem_u32(ring + pos * TC_UNIT), src, bytes, bar_full + 8 * pos);
            src += (size_t)u * TC_UNIT;
            pos += u;
            if (pos >= TC_UNITS) pos = 0;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t par_full = 0, act_par = 0;
      int pos = 0;
      const uint32_t arena_u = smem_u32(arena), ring_u = smem_u32(ring);
      const uint64_t b_const = make_desc_sw128(0, 1024);
      for (int g = blockIdx.x; g < P.n_groups; g += gridDim.x) {
        for (int oi = 0; oi < P.n_ops; ++oi) {
          const TcOp* o = ops_s + oi;
          const int u = o->units, nkb = o->n_kb, kb0 = o->kb_first, N = o->n, nt = o->n_tiles;
          const uint32_t idesc = make_idesc_bf16(128, N);
          const uint64_t a_const = make_desc_sw128(0, (uint32_t)o->sbo);
          uint32_t ts_off[4];
          for (int i = 0; i < 4; ++i) ts_off[i] = (arena_u >> 4) + (uint32_t)(o->tile_slot0[i] * o->slot_stride) * 64u;
          mbar_wait(bar_act, act_par);          // A operand of this op is in shared memory
          act_par ^= 1u;
          tc_fence_after();
          for (int k = 0; k < nkb; ++k) {
            const uint32_t kb = kbs_s[kb0 + k];
            // packed: [0,8) (a_base/1024 + shift) , [8,13) acc_col/16 , [13,16) nk16 , [16] first
            const uint32_t a_slots = kb & 0xFFu, acc_col = ((kb >> 8) & 0x1Fu) << 4, nk16 = (kb >> 13) & 7u;
            uint32_t accum = ((kb >> 16) & 1u) ^ 1u;
            pos = (pos + u - 1) & ~(u - 1);
            if (pos + u > TC_UNITS) pos = 0;
            mbar_wait(bar_full + 8 * pos, (par_full >> pos) & 1u);
            par_full ^= 1u << pos;
            tc_fence_after();
            const uint64_t bd0 = b_const + ((ring_u + pos * TC_UNIT) >> 4);
            for (int mt = 0; mt < nt; ++mt) {
              const uint64_t ad0 = a_const + (ts_off[mt] + a_slots * 64u);
              const uint32_t d_addr = tmem_base + acc_col + mt * N;
              umma_bf16(d_addr, ad0, bd0, idesc, accum);
              for (uint32_t kk = 1; kk < nk16; ++kk) umma_bf16(d_addr, ad0 + 2 * kk, bd0 + 2 * kk, idesc, 1u);
            }
            for (int uu = pos; uu < pos + u; ++uu) umma_commit(bar_empty + 8 * uu);
            pos += u;
            if (pos >= TC_UNITS) pos = 0;
          }
          umma_commit(bar_acc);
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int etid = tid - 64, ew = warp - 2, q = warp & 3, half = ew >> 2;
    const int b = lane & 7, sl = q * 4 + (lane >> 3);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par = 0;
    for (int g = blockIdx.x; g < P.n_groups; g += gridDim.x) {
      const int row = g * TC_G + b;
      const bool row_ok = row < P.R;
      // ---- stage the latent x [8,T,4] fp32 as bf16 hi/lo channels 0..7 of panel 0 (region A, level 0)
      zero_halos(arena, P.zero0_offB, P.zero0_pitch, P.zero0_npanels, etid);
      for (int i = etid; i < P.T * TC_G; i += 256) {
        int t = i >> 3, bb = i & 7, r = g * TC_G + bb;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < P.R) v = reinterpret_cast<const float4*>(P.x)[(size_t)r * P.T + t];
        float hx = __bfloat162float(__float2bfloat16_rn(v.x)), hy = __bfloat162float(__float2bfloat16_rn(v.y));
        float hz = __bfloat162float(__float2bfloat16_rn(v.z)), hw = __bfloat162float(__float2bfloat16_rn(v.w));
        uint4 c0 = make_uint4(pack_bf16(hx, hy), pack_bf16(hz, hw), pack_bf16(v.x - hx, v.y - hy), pack_bf16(v.z - hz, v.w - hw));
        uint8_t* rowp = arena + (t + 2) * 1024 + bb * 128;
        *reinterpret_cast<uint4*>(rowp + ((0 ^ bb) << 4)) = c0;
        *reinterpret_cast<uint4*>(rowp + ((1 ^ bb) << 4)) = make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      epi_bar();
      if (etid == 0) mbar_arrive(bar_act);

      for (int oi = 0; oi < P.n_ops; ++oi) {
        const TcOp* o = ops_s + oi;
        const int epi = o->epi, N = o->n, cout = o->cout, nt = o->n_tiles, n_vt = o->n_vt, t_out = o->t_out;
        const int halfN = N >> 1;
        const bool is_gn = (epi == EPI_GN_TB || epi == EPI_GN_RES_ACC || epi == EPI_GN_RES_ID || epi == EPI_GN);
        // ---- stage per-op parameters while the MMAs run
        {
          const float* pp = P.par + o->par_off;
          for (int i = etid; i < 4 * cout; i += 256) par_s[(i / cout) * 256 + (i % cout)] = pp[i];
          if (epi == EPI_GN_TB) {
            for (int i = etid; i < TC_G * cout; i += 256) {
              int bb = i / cout, c = i - bb * cout, r = g * TC_G + bb;
              float tv = (r < P.R) ? P.tbias[(size_t)r * P.tb_stride + o->tb_off + c] : 0.f;
              if (P.tvec) tv += P.tvec[o->tb_off + c];
              tb_s[bb * TB_LD + c] = tv;
            }
          }
          if (is_gn) for (int i = etid; i < 2048; i += 256) st_s[i] = 0.f;
        }
        epi_bar();
        mbar_wait(bar_acc, acc_par);
        acc_par ^= 1u;
        tc_fence_after();

        // ---- pass 1: GroupNorm statistics over (time, channels of the group) per batch row
        float2 mr[4];   // (mean, rstd) of the 4 groups inside this thread's column half
#pragma unroll
        for (int i = 0; i < 4; ++i) mr[i] = make_float2(0.f, 1.f);
        if (is_gn) {
          for (int vt = 0; vt < n_vt; ++vt) {
            const int lo = o->tile_lo[vt], hi = o->tile_hi[vt];
            if (q * 4 + 4 <= lo || q * 4 >= hi) continue;          // no valid row in this warp's 4 slots
            const bool valid = sl >= lo && sl < hi;
            const int col0 = vt * N + half * halfN;
            for (int ch = 0; ch < halfN; ch += 32) {
              uint32_t r[32];
              tmem_ld32(lane_addr + col0 + ch, r);
              tmem_wait_ld();
              const int c0 = half * halfN + ch;
#pragma unroll
              for (int sb = 0; sb < 4; ++sb) {
                const float4 b0 = *reinterpret_cast<const float4*>(par_s + c0 + sb * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(par_s + c0 + sb * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float s = 0.f, ss = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float v = __uint_as_float(r[sb * 8 + j]) + bb[j];
                  s += v; ss = fmaf(v, v, ss);
                }
                if (!valid) { s = 0.f; ss = 0.f; }
                s += __shfl_xor_sync(0xffffffffu, s, 8);  ss += __shfl_xor_sync(0xffffffffu, ss, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16); ss += __shfl_xor_sync(0xffffffffu, ss, 16);
                if (lane < 8) {
                  // slot owned by (quadrant, b, sub-block): only this lane of this warp touches it -> deterministic
                  float2* sp = reinterpret_cast<float2*>(&st_s[((q * 8 + b) * 32 + (c0 >> 3) + sb) * 2]);
                  float2 cur = *sp;
                  cur.x += s; cur.y += ss;
                  *sp = cur;
                }
              }
            }
          }
          epi_bar();
          // (row, group) mean / rstd: 64 threads, fixed summation order
          if (etid < 64) {
            const int bb = etid >> 3, gg = etid & 7, sbpg = o->cpg >> 3;
            float S = 0.f, SS = 0.f;
            for (int k = 0; k < sbpg; ++k)
              for (int qq = 0; qq < 4; ++qq) {
                S += st_s[((qq * 8 + bb) * 32 + gg * sbpg + k) * 2];
                SS += st_s[((qq * 8 + bb) * 32 + gg * sbpg + k) * 2 + 1];
              }
            const float inv_n = 1.0f / (float)(t_out * o->cpg);
            const float mean = S * inv_n;
            const float var = fmaxf(SS * inv_n - mean * mean, 0.f);
            mr_s[(bb * 8 + gg) * 2] = mean;
            mr_s[(bb * 8 + gg) * 2 + 1] = rsqrtf(var + 1e-5f);
          }
          epi_bar();
#pragma unroll
          for (int i = 0; i < 4; ++i) mr[i] = *reinterpret_cast<const float2*>(&mr_s[(b * 8 + half * 4 + i) * 2]);
        }

        // ---- pass 2: normalise / activate / add, write the next A operand (or eps)
        const int cpg_shift = 31 - __clz(o->cpg);
        const bool dbg = (o->dbg_stage >= 0 && o->dbg_stage == P.dbg_stage && P.dbg_out != nullptr);
        if (epi == EPI_OUT) {
          if (half == 0) {
            for (int vt = 0; vt < n_vt; ++vt) {
              uint32_t r[16];
              tmem_ld16(lane_addr + vt * N, r);
              tmem_wait_ld();
              const int slot = o->tile_slot0[vt] + sl;
              if (sl >= o->tile_lo[vt] && sl < o->tile_hi[vt] && row_ok) {
                float4 v = make_float4(__uint_as_float(r[0]) + par_s[0], __uint_as_float(r[1]) + par_s[1],
                                       __uint_as_float(r[2]) + par_s[2], __uint_as_float(r[3]) + par_s[3]);
                reinterpret_cast<float4*>(P.eps)[(size_t)row * P.T + slot] = v;
              }
            }
          }
        } else {
          for (int vt = 0; vt < n_vt; ++vt) {
            const int mt = vt % nt, ph = vt / nt;
            const int lo = o->tile_lo[mt], hi = o->tile_hi[mt];
            if (q * 4 + 4 <= lo || q * 4 >= hi) continue;
            const bool valid = sl >= lo && sl < hi;
            const int slot = (epi == EPI_UP) ? 2 * (o->tile_slot0[mt] + sl) + ph : o->tile_slot0[mt] + sl;
            const int col0 = vt * N + half * halfN;
            uint8_t* rowp = arena + o->dst_off + (slot + 2) * 1024 + b * 128;
            for (int ch = 0; ch < halfN; ch += 32) {
              uint32_t r[32], rr[32];
              tmem_ld32(lane_addr + col0 + ch, r);
              if (epi == EPI_GN_RES_ACC) tmem_ld32(lane_addr + o->res_col + col0 + ch, rr);
              tmem_wait_ld();
              const int c0 = half * halfN + ch;
#pragma unroll
              for (int sb = 0; sb < 4; ++sb) {
                const int c = c0 + sb * 8;
                const float2 m2 = mr[((c >> cpg_shift) & 3)];
                uint8_t* dstp = rowp + (c >> 6) * o->dst_pitch + ((((c >> 3) & 7) ^ b) << 4);
                float y[8];
                {
                  const float4 b0 = *reinterpret_cast<const float4*>(par_s + c), b1 = *reinterpret_cast<const float4*>(par_s + c + 4);
                  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(r[sb * 8 + j]) + bb[j];
                }
                if (is_gn) {
                  const float4 g0 = *reinterpret_cast<const float4*>(par_s + 256 + c), g1 = *reinterpret_cast<const float4*>(par_s + 256 + c + 4);
                  const float4 e0 = *reinterpret_cast<const float4*>(par_s + 512 + c), e1 = *reinterpret_cast<const float4*>(par_s + 512 + c + 4);
                  const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                  const float bt[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float sc = m2.y * gm[j];
                    y[j] = mish_fast(fmaf(y[j] - m2.x, sc, bt[j]));
                  }
                }
                if (epi == EPI_GN_TB) {
                  const float4 t0 = *reinterpret_cast<const float4*>(tb_s + b * TB_LD + c), t1 = *reinterpret_cast<const float4*>(tb_s + b * TB_LD + c + 4);
                  y[0] += t0.x; y[1] += t0.y; y[2] += t0.z; y[3] += t0.w; y[4] += t1.x; y[5] += t1.y; y[6] += t1.z; y[7] += t1.w;
                } else if (epi == EPI_GN_RES_ACC) {
                  const float4 s0 = *reinterpret_cast<const float4*>(par_s + 768 + c), s1 = *reinterpret_cast<const float4*>(par_s + 768 + c + 4);
                  const float rb[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                  for (int j = 0; j < 8; ++j) y[j] += __uint_as_float(rr[sb * 8 + j]) + rb[j];
                } else if (epi == EPI_GN_RES_ID) {
                  if (valid) {
                    uint4 old = *reinterpret_cast<const uint4*>(dstp);
                    float2 f0 = unpack_bf16(old.x), f1 = unpack_bf16(old.y), f2 = unpack_bf16(old.z), f3 = unpack_bf16(old.w);
                    y[0] += f0.x; y[1] += f0.y; y[2] += f1.x; y[3] += f1.y;
                    y[4] += f2.x; y[5] += f2.y; y[6] += f3.x; y[7] += f3.y;
                  }
                }
                if (valid && c < cout) {
                  uint4 pk = make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
                  *reinterpret_cast<uint4*>(dstp) = pk;
                  if (o->save_skip >= 0) {
                    uint8_t* gp = P.skipbuf + (size_t)blockIdx.x * P.skip_stride + o->save_skip +
                                  ((size_t)(c >> 6) * t_out + slot) * 1024 + b * 128 + ((((c >> 3) & 7) ^ b) << 4);
                    *reinterpret_cast<uint4*>(gp) = pk;
                  }
                  if (dbg && row_ok) {
                    float* dp = P.dbg_out + ((size_t)row * t_out + slot) * cout + c;
                    float2 f0 = unpack_bf16(pk.x), f1 = unpack_bf16(pk.y), f2 = unpack_bf16(pk.z), f3 = unpack_bf16(pk.w);
                    dp[0] = f0.x; dp[1] = f0.y; dp[2] = f1.x; dp[3] = f1.y; dp[4] = f2.x; dp[5] = f2.y; dp[6] = f3.x; dp[7] = f3.y;
                  }
                }
              }
            }
          }
        }
        // ---- level change: zero the halo slots of the new layout; reload a skip connection
        if (o->zero_pitch) zero_halos(arena, o->zero_offB, o->zero_pitch, o->zero_npanels, etid);
        if (o->load_skip >= 0) {
          epi_bar();     // all skip stores of this CTA are older than this point; make them visible
          const uint8_t* gp = P.skipbuf + (size_t)blockIdx.x * P.skip_stride + o->load_skip;
          const int per_panel = o->load_T * 64;          // uint4 per panel
          for (int i = etid; i < o->load_npanels * per_panel; i += 256) {
            int p = i / per_panel, w = i - p * per_panel;
            *reinterpret_cast<uint4*>(arena + o->load_off + p * o->load_pitch + 2048 + w * 16) =
                *reinterpret_cast<const uint4*>(gp + (size_t)i * 16);
          }
        }
        tc_fence_before();
        fence_proxy_async();
        epi_bar();
        if (etid == 0 && oi + 1 < P.n_ops) mbar_arrive(bar_act);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// weight packing: one k-block = [rows][64 k] bf16, 128B-swizzled image ready for a flat bulk copy
// ------------------------------------------------------------------------------------------------
__global__ void tc_pack_tile_kernel(uint8_t* __restrict__ dst, const float* __restrict__ w, int cout, int cin, int K,
                                    int transposed, int tap, int ci0, int rows, int dup4) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 64) return;
  int n = idx >> 6, k = idx & 63;
  int ci = ci0 + k;
  if (dup4) ci = (k < 8) ? (k & 3) : cin;     // first layer: channels 0..3 = hi part, 4..7 = lo part of x
  float v = 0.f;
  if (n < cout && ci < cin) v = transposed ? w[((size_t)ci * cout + n) * K + tap] : w[((size_t)n * cin + ci) * K + tap];
  __nv_bfloat16 hv = __float2bfloat16_rn(v);
  *reinterpret_cast<__nv_bfloat16*>(dst + sw128_off(n, k >> 3) + (k & 7) * 2) = hv;
}

static TcState* st_of(CldHandle* h) { return reinterpret_cast<TcState*>(h->tc); }

bool tc_enabled(const CldHandle* h) {
  const CldConfig& c = h->cfg;
  return c.precision == CLD_PREC_BF16 && c.dims[0] == 64 && c.dims[1] == 128 && c.dims[2] == 256 && c.horizon <= 64 &&
         c.horizon >= 16 && c.latent_dim == 4;
}

int tc_pack_block(CldHandle* h, int exec_idx, int cin, int cout, const float* c0w, const float* c0b, const float* g0,
                  const float* b0, const float* c1w, const float* c1b, const float* g1, const float* b1, const float* rw,
                  const float* rb, cudaStream_t) {
  if (!h->tc) h->tc = new TcState();
  st_of(h)->blk[exec_idx] = {cin, cout, c0w, c0b, g0, b0, c1w, c1b, g1, b1, rw, rb};
  return 0;
}
int tc_pack_down(CldHandle* h, int lvl, int ch, const float* w, const float* b, cudaStream_t) {
  if (!h->tc) h->tc = new TcState();
  st_of(h)->down[lvl] = {ch, w, b};
  return 0;
}
int tc_pack_up(CldHandle* h, int lvl, int ch, const float* w, const float* b, cudaStream_t) {
  if (!h->tc) h->tc = new TcState();
  st_of(h)->up[lvl] = {ch, w, b};
  return 0;
}
int tc_pack_final(CldHandle* h, const float* fw, const float* fb, const float* fg, const float* fbt, const float* f1w,
                  const float* f1b, cudaStream_t) {
  if (!h->tc) h->tc = new TcState();
  TcState* s = st_of(h);
  s->fw = fw; s->fb = fb; s->fg = fg; s->fbt = fbt; s->f1w = f1w; s->f1b = f1b;
  return 0;
}

namespace {
struct Level { int T, pitch, offB, npanels, n_tiles, slot0[4], lo[4], hi[4]; };

struct Builder {
  CldHandle* h; TcState* s; cudaStream_t stream;
  std::vector<uint8_t*> tile_dst;    // unused
  size_t w_bytes = 0; size_t par_floats = 0;
  struct PackJob { size_t off; const float* w; int cout, cin, K, transposed, tap, ci0, rows, dup4; };
  struct ParJob { size_t off; const float* src; int n; };
  std::vector<PackJob> packs; std::vector<ParJob> pars;

  int add_tile(const float* w, int cout, int cin, int K, int transposed, int tap, int ci0, int rows, int dup4, int units) {
    size_t off = w_bytes;
    packs.push_back({off, w, cout, cin, K, transposed, tap, ci0, rows, dup4});
    w_bytes += (size_t)units * TC_UNIT;
    return (int)off;
  }
  int add_par(const float* bias, const float* gamma, const float* beta, const float* resb, int cout) {
    size_t off = par_floats;
    const float* srcs[4] = {bias, gamma, beta, resb};
    for (int i = 0; i < 4; ++i) if (srcs[i]) pars.push_back({off + (size_t)i * cout, srcs[i], cout});
    par_floats += 4 * (size_t)cout;
    return (int)off;
  }
};

void set_tiles(TcOp& o, const Level& L) {
  o.n_tiles = L.n_tiles;
  for (int i = 0; i < 4; ++i) { o.tile_slot0[i] = L.slot0[i]; o.tile_lo[i] = L.lo[i]; o.tile_hi[i] = L.hi[i]; }
}
TcOp blank_op() {
  TcOp o;
  memset(&o, 0, sizeof(o));
  o.sbo = 1024; o.slot_stride = 1; o.save_skip = -1; o.load_skip = -1; o.dbg_stage = -1; o.cpg = 8;
  return o;
}
}  // namespace

int tc_finalize(CldHandle* h, cudaStream_t stream) {
  TcState* s = st_of(h);
  if (!s) return fail(h, CLD_ERR_STATE, "tc_finalize without weights");
  const CldConfig& c = h->cfg;
  const int T = c.horizon;
  Level L[3];
  const int np[3] = {1, 2, 4};
  for (int l = 0; l < 3; ++l) {
    Level& v = L[l];
    v.T = T >> l; v.pitch = (v.T + 2) * 1024; v.npanels = np[l]; v.offB = np[l] * v.pitch + 2048;
    v.n_tiles = (v.T + 15) / 16;
    for (int i = 0; i < 4; ++i) { v.slot0[i] = 0; v.lo[i] = 0; v.hi[i] = 0; }
    for (int i = 0; i < v.n_tiles; ++i) {
      bool last = i == v.n_tiles - 1;
      v.slot0[i] = (last && v.T >= 16) ? v.T - 16 : 16 * i;
      v.lo[i] = 16 * i - v.slot0[i];
      v.hi[i] = (v.T - v.slot0[i] < 16) ? v.T - v.slot0[i] : 16;
    }
    // every 16-slot tile read (plus taps, plus stride-2 reads from the level above) must stay inside the arena
    if (v.offB + v.npanels * v.pitch + 5 * 1024 > TC_ARENA) return fail(h, CLD_ERR_UNSUPPORTED, "horizon too long for the bf16 arena");
  }
  Builder B{h, s, stream};
  s->ops.clear(); s->kbs.clear();
  const int tb_off_exec[12] = {0, 64, 128, 256, 384, 640, 896, 1152, 1408, 1536, 1664, 1728};
  auto units_of = [](int n) { int b = n * 128; return b <= TC_UNIT ? 1 : b / TC_UNIT; };

  // generic conv accumulation: taps (k index, input slot offset) x input panels
  auto add_conv = [&](TcOp& o, const float* w, int cout, int cin, int K, int transposed, const int* tap_k, const int* tap_off,
                      int ntaps, const int* panel_base, int npanels_in, int acc_col, int n_rows, bool dup4, int stride) {
    bool first = true;
    for (int t = 0; t < ntaps; ++t)
      for (int p = 0; p < npanels_in; ++p) {
        TcKb kb;
        kb.a_base = panel_base[p];
        kb.shift = (stride == 2) ? tap_off[t] + 2 : tap_off[t] + 2;
        kb.w_off = B.add_tile(w, cout, cin, K, transposed, tap_k[t], p * 64, n_rows, dup4 ? 1 : 0, o.units);
        kb.acc_col = acc_col; kb.nk16 = dup4 ? 1 : ((cin - p * 64 >= 64) ? 4 : (cin - p * 64 + 15) / 16);
        kb.first = first ? 1 : 0;
        first = false;
        s->kbs.push_back(kb);
        o.n_kb++;
      }
  };
  const int k5[5] = {0, 1, 2, 3, 4}, o5[5] = {-2, -1, 0, 1, 2}, k1[1] = {0}, o1[1] = {0};
  const int k3[3] = {0, 1, 2}, o3[3] = {-1, 0, 1};
  const int kue[2] = {1, 3}, oue[2] = {0, -1}, kuo[2] = {0, 2}, ouo[2] = {1, 0};

  auto res_block = [&](int e, int lvl, const int* in_panels, int n_in, bool concat, int stage) {
    const TcState::Blk& bk = s->blk[e];
    const Level& lv = L[lvl];
    const int cout = bk.cout;
    const bool has_res = bk.rw != nullptr;
    // ---- op A: conv0 (+ residual 1x1 conv into the second accumulator set)
    TcOp a = blank_op();
    a.n = cout; set_tiles(a, lv); a.units = units_of(cout); a.kb_bytes = cout * 128; a.kb_first = (int)s->kbs.size();
    a.epi = EPI_GN_TB; a.cout = cout; a.cpg = cout / 8; a.t_out = lv.T; a.n_vt = lv.n_tiles;
    a.dst_off = concat ? 0 : lv.offB; a.dst_pitch = lv.pitch;
    a.par_off = B.add_par(bk.c0b, bk.g0, bk.b0, nullptr, cout); a.tb_off = tb_off_exec[e]; a.res_col = 256;
    add_conv(a, bk.c0w, cout, bk.cin, 5, 0, k5, o5, 5, in_panels, n_in, 0, cout, e == 0, 1);
    if (has_res) add_conv(a, bk.rw, cout, bk.cin, 1, 0, k1, o1, 1, in_panels, n_in, 256, cout, e == 0, 1);
    s->ops.push_back(a);
    // ---- op B: conv1, GroupNorm + Mish + residual
    TcOp b = blank_op();
    b.n = cout; set_tiles(b, lv); b.units = units_of(cout); b.kb_bytes = cout * 128; b.kb_first = (int)s->kbs.size();
    b.epi = has_res ? EPI_GN_RES_ACC : EPI_GN_RES_ID; b.cout = cout; b.cpg = cout / 8; b.t_out = lv.T; b.n_vt = lv.n_tiles;
    b.dst_off = 0; b.dst_pitch = lv.pitch; b.res_col = 256;
    b.par_off = B.add_par(bk.c1b, bk.g1, bk.b1, bk.rb, cout);
    int hp[4];
    for (int p = 0; p < cout / 64; ++p) hp[p] = a.dst_off + p * lv.pitch;
    add_conv(b, bk.c1w, cout, cout, 5, 0, k5, o5, 5, hp, cout / 64, 0, cout, false, 1);
    b.dbg_stage = stage;
    s->ops.push_back(b);
  };
  auto panels_of = [&](int off, int pitch, int n, int* out) { for (int p = 0; p < n; ++p) out[p] = off + p * pitch; };
  const int skip1_off = 0, skip1_bytes = 2 * L[1].T * 1024, skip2_off = skip1_bytes, skip2_bytes = 4 * L[2].T * 1024;
  int pa[8];

  // level 0
  panels_of(0, L[0].pitch, 1, pa); res_block(0, 0, pa, 1, false, 0);
  res_block(1, 0, pa, 1, false, 1);
  {  // downs.0.2: k3 stride 2, 64 -> 64, output at level 1
    TcOp d = blank_op();
    d.n = 64; set_tiles(d, L[1]); d.sbo = 2048; d.slot_stride = 2; d.units = 1; d.kb_bytes = 64 * 128; d.kb_first = (int)s->kbs.size();
    d.epi = EPI_BIAS; d.cout = 64; d.t_out = L[1].T; d.n_vt = L[1].n_tiles; d.dst_off = 0; d.dst_pitch = L[1].pitch;
    d.par_off = B.add_par(s->down[0].b, nullptr, nullptr, nullptr, 64);
    d.zero_pitch = L[1].pitch; d.zero_npanels = L[1].npanels; d.zero_offB = L[1].offB; d.dbg_stage = 2;
    const int o3s[3] = {-1, 0, 1};
    add_conv(d, s->down[0].w, 64, 64, 3, 0, k3, o3s, 3, pa, 1, 0, 64, false, 2);
    s->ops.push_back(d);
  }
  // level 1
  panels_of(0, L[1].pitch, 1, pa); res_block(2, 1, pa, 1, false, 3);
  panels_of(0, L[1].pitch, 2, pa); res_block(3, 1, pa, 2, false, 4);
  s->ops.back().save_skip = skip1_off;
  {  // downs.1.2
    TcOp d = blank_op();
    d.n = 128; set_tiles(d, L[2]); d.sbo = 2048; d.slot_stride = 2; d.units = 2; d.kb_bytes = 128 * 128; d.kb_first = (int)s->kbs.size();
    d.epi = EPI_BIAS; d.cout = 128; d.t_out = L[2].T; d.n_vt = L[2].n_tiles; d.dst_off = 0; d.dst_pitch = L[2].pitch;
    d.par_off = B.add_par(s->down[1].b, nullptr, nullptr, nullptr, 128);
    d.zero_pitch = L[2].pitch; d.zero_npanels = L[2].npanels; d.zero_offB = L[2].offB; d.dbg_stage = 5;
    add_conv(d, s->down[1].w, 128, 128, 3, 0, k3, o3, 3, pa, 2, 0, 128, false, 2);
    s->ops.push_back(d);
  }
  // level 2
  panels_of(0, L[2].pitch, 2, pa); res_block(4, 2, pa, 2, false, 6);
  panels_of(0, L[2].pitch, 4, pa); res_block(5, 2, pa, 4, false, 7);
  s->ops.back().save_skip = skip2_off;
  res_block(6, 2, pa, 4, false, 8);
  res_block(7, 2, pa, 4, false, 9);
  {  // after mid_block2: bring skip2 back into region B
    TcOp& o = s->ops.back();
    o.load_skip = skip2_off; o.load_off = L[2].offB; o.load_pitch = L[2].pitch; o.load_npanels = 4; o.load_T = L[2].T;
  }
  panels_of(0, L[2].pitch, 4, pa); panels_of(L[2].offB, L[2].pitch, 4, pa + 4);
  res_block(8, 2, pa, 8, true, 10);
  panels_of(0, L[2].pitch, 2, pa); res_block(9, 2, pa, 2, false, 11);
  {  // ups.0.2: transposed conv 128 -> 128, level 2 -> level 1 (two output phases)
    TcOp u = blank_op();
    u.n = 128; set_tiles(u, L[2]); u.units = 2; u.kb_bytes = 128 * 128; u.kb_first = (int)s->kbs.size();
    u.epi = EPI_UP; u.cout = 128; u.t_out = L[1].T; u.n_vt = 2 * L[2].n_tiles; u.dst_off = 0; u.dst_pitch = L[1].pitch;
    u.par_off = B.add_par(s->up[0].b, nullptr, nullptr, nullptr, 128);
    u.zero_pitch = L[1].pitch; u.zero_npanels = L[1].npanels; u.zero_offB = L[1].offB; u.dbg_stage = 12;
    u.load_skip = skip1_off; u.load_off = L[1].offB; u.load_pitch = L[1].pitch; u.load_npanels = 2; u.load_T = L[1].T;
    add_conv(u, s->up[0].w, 128, 128, 4, 1, kue, oue, 2, pa, 2, 0, 128, false, 1);
    add_conv(u, s->up[0].w, 128, 128, 4, 1, kuo, ouo, 2, pa, 2, L[2].n_tiles * 128, 128, false, 1);
    s->ops.push_back(u);
  }
  // level 1 (up path)
  panels_of(0, L[1].pitch, 2, pa); panels_of(L[1].offB, L[1].pitch, 2, pa + 2);
  res_block(10, 1, pa, 4, true, 13);
  panels_of(0, L[1].pitch, 1, pa); res_block(11, 1, pa, 1, false, 14);
  {  // ups.1.2: transposed conv 64 -> 64, level 1 -> level 0
    TcOp u = blank_op();
    u.n = 64; set_tiles(u, L[1]); u.units = 1; u.kb_bytes = 64 * 128; u.kb_first = (int)s->kbs.size();
    u.epi = EPI_UP; u.cout = 64; u.t_out = L[0].T; u.n_vt = 2 * L[1].n_tiles; u.dst_off = 0; u.dst_pitch = L[0].pitch;
    u.par_off = B.add_par(s->up[1].b, nullptr, nullptr, nullptr, 64);
    u.zero_pitch = L[0].pitch; u.zero_npanels = L[0].npanels; u.zero_offB = L[0].offB; u.dbg_stage = 15;
    add_conv(u, s->up[1].w, 64, 64, 4, 1, kue, oue, 2, pa, 1, 0, 64, false, 1);
    add_conv(u, s->up[1].w, 64, 64, 4, 1, kuo, ouo, 2, pa, 1, L[1].n_tiles * 64, 64, false, 1);
    s->ops.push_back(u);
  }
  {  // final_conv.0: conv k5 + GroupNorm + Mish -> region B ; final_conv.1: 1x1 conv 64 -> 4 -> eps
    panels_of(0, L[0].pitch, 1, pa);
    TcOp f = blank_op();
    f.n = 64; set_tiles(f, L[0]); f.units = 1; f.kb_bytes = 64 * 128; f.kb_first = (int)s->kbs.size();
    f.epi = EPI_GN; f.cout = 64; f.cpg = 8; f.t_out = L[0].T; f.n_vt = L[0].n_tiles; f.dst_off = L[0].offB; f.dst_pitch = L[0].pitch;
    f.par_off = B.add_par(s->fb, s->fg, s->fbt, nullptr, 64); f.dbg_stage = 16;
    add_conv(f, s->fw, 64, 64, 5, 0, k5, o5, 5, pa, 1, 0, 64, false, 1);
    s->ops.push_back(f);
    TcOp g = blank_op();
    g.n = 16; set_tiles(g, L[0]); g.units = 1; g.kb_bytes = 16 * 128; g.kb_first = (int)s->kbs.size();
    g.epi = EPI_OUT; g.cout = 4; g.t_out = L[0].T; g.n_vt = L[0].n_tiles;
    g.par_off = B.add_par(s->f1b, nullptr, nullptr, nullptr, 4);
    panels_of(L[0].offB, L[0].pitch, 1, pa);
    add_conv(g, s->f1w, 4, 64, 1, 0, k1, o1, 1, pa, 1, 0, 16, false, 1);
    s->ops.push_back(g);
  }
  // the EPI_UP / res accumulators must fit the 512 TMEM columns
  for (const TcOp& o : s->ops) {
    int cols = o.n_vt * o.n + (o.epi == EPI_GN_TB ? 256 : 0);
    if (o.n_vt * o.n > 256 || cols > 512) return fail(h, CLD_ERR_UNSUPPORTED, "accumulators exceed tensor memory");
  }

  // ---- materialise blobs on the device
  auto alloc = [&](void** p, size_t bytes) -> int {
    CLD_CUDA_OK(h, cudaMalloc(p, bytes));
    h->allocs.push_back(*p);
    return 0;
  };
  int rc;
  s->wblob_bytes = B.w_bytes; s->par_floats = B.par_floats;
  if ((rc = alloc((void**)&s->wblob, B.w_bytes))) return rc;
  if ((rc = alloc((void**)&s->par, B.par_floats * sizeof(float)))) return rc;
  CLD_CUDA_OK(h, cudaMemsetAsync(s->wblob, 0, B.w_bytes, stream));
  CLD_CUDA_OK(h, cudaMemsetAsync(s->par, 0, B.par_floats * sizeof(float), stream));
  for (const auto& j : B.packs) {
    tc_pack_tile_kernel<<<(j.rows * 64 + 255) / 256, 256, 0, stream>>>(s->wblob + j.off, j.w, j.cout, j.cin, j.K, j.transposed,
                                                                      j.tap, j.ci0, j.rows, j.dup4);
  }
  CLD_LAUNCH_OK(h, "tc_pack_tile_kernel");
  for (const auto& j : B.pars)
    CLD_CUDA_OK(h, cudaMemcpyAsync(s->par + j.off, j.src, j.n * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  if ((rc = alloc((void**)&s->d_ops, s->ops.size() * sizeof(TcOp)))) return rc;
  if (s->ops.size() > (size_t)TC_MAX_OPS || s->kbs.size() > (size_t)TC_MAX_KBS)
    return fail(h, CLD_ERR_UNSUPPORTED, "op table too large (%zu ops, %zu k-blocks)", s->ops.size(), s->kbs.size());
  std::vector<uint32_t> packed(s->kbs.size());
  for (TcOp& o : s->ops) {
    o.w_first = s->kbs[o.kb_first].w_off;
    for (int k = 0; k < o.n_kb; ++k) {
      const TcKb& kb = s->kbs[o.kb_first + k];
      if (kb.w_off != o.w_first + k * o.units * TC_UNIT || kb.a_base % 1024 || kb.acc_col % 16)
        return fail(h, CLD_ERR_UNSUPPORTED, "internal: k-block record not packable");
      uint32_t a_slots = (uint32_t)(kb.a_base / 1024 + kb.shift);
      if (a_slots > 255u || kb.acc_col > 496 || kb.nk16 > 4) return fail(h, CLD_ERR_UNSUPPORTED, "internal: k-block field overflow");
      packed[o.kb_first + k] = a_slots | ((uint32_t)(kb.acc_col / 16) << 8) | ((uint32_t)kb.nk16 << 13) | ((uint32_t)kb.first << 16);
    }
  }
  if ((rc = alloc((void**)&s->d_kbs, packed.size() * sizeof(uint32_t)))) return rc;
  CLD_CUDA_OK(h, cudaMemcpyAsync(s->d_ops, s->ops.data(), s->ops.size() * sizeof(TcOp), cudaMemcpyHostToDevice, stream));
  CLD_CUDA_OK(h, cudaMemcpyAsync(s->d_kbs, packed.data(), packed.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  CLD_CUDA_OK(h, cudaStreamSynchronize(stream));
  s->grid = h->num_sms;
  s->skip_stride = ((skip1_bytes + skip2_bytes + 1023) / 1024) * 1024;
  if ((rc = alloc((void**)&s->skipbuf, (size_t)s->grid * s->skip_stride))) return rc;
  CLD_CUDA_OK(h, cudaFuncSetAttribute(unet_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
  CLD_CUDA_OK(h, cudaStreamSynchronize(stream));
  s->ready = true;
  (void)skip2_off;
  return 0;
}

// time / cond projection shared with the fp32 path (kernels_unet_fp32.cu)
int unet_time_bias(CldHandle* h, const float* cond, const int64_t* t, int R, cudaStream_t s);

static int tc_launch(CldHandle* h, const float* x, float* eps, int R, const float* tvec, cudaStream_t stream);

int tc_unet_forward(CldHandle* h, const float* x, const float* cond, const int64_t* t, float* eps, int R, cudaStream_t stream) {
  TcState* s = st_of(h);
  if (!s || !s->ready) return fail(h, CLD_ERR_STATE, "bf16 denoiser weights not packed");
  int rc;
  if ((rc = unet_time_bias(h, cond, t, R, stream))) return rc;
  return tc_launch(h, x, eps, R, nullptr, stream);
}

int tc_unet_forward_prepared(CldHandle* h, const float* x, float* eps, int R, cudaStream_t stream) {
  TcState* s = st_of(h);
  if (!s || !s->ready) return fail(h, CLD_ERR_STATE, "bf16 denoiser weights not packed");
  return tc_launch(h, x, eps, R, h->tvec, stream);
}

static int tc_launch(CldHandle* h, const float* x, float* eps, int R, const float* tvec, cudaStream_t stream) {
  TcState* s = st_of(h);
  TcParams P;
  P.tvec = tvec;
  P.ops = s->d_ops; P.n_ops = (int)s->ops.size(); P.kbs = s->d_kbs; P.n_kbs = (int)s->kbs.size(); P.wblob = s->wblob; P.par = s->par;
  P.tbias = h->tbias; P.tb_stride = h->unet.tb_total; P.x = x; P.eps = eps; P.R = R; P.T = h->cfg.horizon;
  P.n_groups = (R + TC_G - 1) / TC_G; P.skipbuf = s->skipbuf; P.skip_stride = s->skip_stride;
  const int T = h->cfg.horizon;
  P.zero0_pitch = (T + 2) * 1024; P.zero0_npanels = 1; P.zero0_offB = (T + 2) * 1024 + 2048;
  P.dbg_stage = h->dbg_out ? h->dbg_stage : -1; P.dbg_out = h->dbg_out;
  int grid = P.n_groups < s->grid ? P.n_groups : s->grid;
  unet_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(P);
  CLD_LAUNCH_OK(h, "unet_tc_kernel");
  return 0;
}

void tc_destroy(CldHandle* h) {
  if (h->tc) { delete st_of(h); h->tc = nullptr; }
}

}  // namespace cld
