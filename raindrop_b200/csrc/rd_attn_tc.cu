// Temporal self-attention for short sequences (T <= 64, head_dim <= 96) on the 5th-gen tensor cores.
//
// nn.TransformerEncoder's attention as called at code/models_rd.py:358 for the P19 shape (T = 60, hd = 76):
// per (sample, head) S = scale * Q K^T, key-padding-masked softmax, attention dropout, O = P V -- and the whole
// backward (dV = Pd^T dO, dP = dO V^T, dS = P * (dP - rowsum(dP * P)), dQ = scale * dS K, dK = scale * dS^T Q).
// One CTA per (sample, head); every contraction is a tcgen05.mma.kind::tf32 with error compensation (operands split
// as hi + lo, three MMAs per k-step: lo.hi + hi.lo + hi.hi), so the results are fp32-accurate like the CUDA-core
// kernels they replace (rd_attn_small.cu).  Nothing T x T ever reaches HBM; the backward RECOMPUTES the
// probabilities from Q, K and the counter-based dropout stream.
//
// Data movement: each [T x hd] head slice of the packed qkv / d(ctx) tensors is fetched by TMA as three
// {32 column, 64 row} boxes (5-D tensor map over [T, B, 3, H, hd]: out-of-range columns >= hd and rows >= T arrive
// as zeros).  An operand is used in one of two roles and each role has its own shared-memory image:
//   * K-major   (contraction over the COLUMNS d):  S = Q K^T, dP = dO V^T   -- classic 128B swizzle (16-byte atoms)
//   * MN-major  (contraction over the ROWS t):     O = P V, dV = Pd^T dO, dQ = dS K, dK = dS^T Q
//                                                  -- 128B swizzle with 32-BYTE atoms, the only MN-major layout the
//                                                     tensor core takes for 32-bit operands (rd_tc_common.cuh)
// The same global tile is simply fetched through a second tensor map when the other role is needed.  The
// probabilities / score gradients are written by the softmax threads as row-major [64 x 64] tiles in whichever
// swizzle their consumer needs: read row-wise they are a K-major A operand (O = P V, dQ = dS K), read column-wise
// an MN-major A operand (dV = Pd^T dO, dK = dS^T Q) -- no transposes anywhere.  M = 128 MMAs are issued on 64-row
// tiles: rows 64..127 of the A operand read whatever follows in shared memory and only produce accumulator rows
// nobody reads.
//
// Backward, shared-memory plan (four 48 KB regions, 200 KB with the tail pad):
//   phase 1  R0 = Q, R1 = K, R2 = V, R3 = dO (K-major images)          S = Q K^T, dP = dO V^T
//   phase 2  softmax threads: R0 <- Pd (MN image), R1 <- scale*dS (K-major image), R0/R1 tails <- scale*dS (MN image)
//            TMA meanwhile:   R2 <- dO (MN image), R3 <- K (MN image)   dV = Pd^T dO, dQ = dS K
//   phase 3  R2 <- Q (MN image)                                          dK = dS^T Q
#include <stdlib.h>

#include "rd_kernels.cuh"
#include "rd_tc_common.cuh"

namespace rd {
using namespace tc;
namespace {

constexpr int TR = 64;                    // rows (timestamps) per tile
constexpr int NG = 3;                     // 32-column groups per head slice (hd <= 96)
constexpr int GRP = TR * 128;             // one {32 col, 64 row} box: 8192 bytes
constexpr int TILE = NG * GRP;            // 24576 bytes (hi); the lo image follows
constexpr int PT = 2 * GRP;               // probability / score-gradient tile [64 x 64]: 16384 bytes
constexpr int NTHR = 128;

struct AttnTcP {
  float* ctx; float* dqkv;
  const int64_t* lengths;
  int B, H, T, hd, D;
  float scale, drop_p;
  const uint64_t* rng; uint32_t site;
  unsigned long long* dbg;      // optional phase timestamps (rd_debug_attention_timing): [CTA][16] of %globaltimer
};

__device__ __forceinline__ void stamp(const AttnTcP& p, int slot) {
  if (p.dbg && threadIdx.x == 0) {
    // SM cycle counter: a %globaltimer read costs 0.25-0.5 us (and ticks every 256 ns), which perturbed what it measured
    p.dbg[(size_t)blockIdx.x * 16 + slot] = (unsigned long long)clock64();
  }
}

__device__ __forceinline__ void stamp_by(const AttnTcP& p, int slot, int tid) {
  if (p.dbg && (int)threadIdx.x == tid) p.dbg[(size_t)blockIdx.x * 16 + slot] = (unsigned long long)clock64();
}
__device__ __forceinline__ float lo_of(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// lo image (at +lo_off) of `bytes` bytes of hi image; same addresses, so the swizzle never has to be undone
__device__ __forceinline__ void lo_pass(uint32_t hi, uint32_t lo_off, uint32_t bytes) {
  lo_image<6>(hi, hi + lo_off, bytes / 16u, threadIdx.x, (uint32_t)NTHR);
}

// D[128 x N] (+)= A . B with error compensation.  Operand images: hi at base, lo at base + *_lo.
//   KMAJOR operand: contraction over columns; k-step ks covers columns 8*ks..8*ks+7 (group ks/4, 32 bytes * (ks%4))
//   MN operand:     contraction over rows;    k-step ks covers rows 8*ks..8*ks+7 (1024 bytes apart), groups GRP apart,
//                   image in the 32-byte-atom swizzle
// The four base descriptors are built once; a k-step only moves the 14-bit start-address field (units of 16 bytes), so
// the single issuing thread spends a handful of instructions per MMA (building descriptors inside the loop made the
// issue of 36 MMAs take 3 us -- one thread's dependent 64-bit arithmetic, not the tensor pipe).
__device__ __forceinline__ void umma_tf32_split(uint32_t d_tmem, uint32_t a_lo32, uint32_t a_hi32, uint32_t b_lo32, uint32_t b_hi32,
                                                uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_lo32), "r"(a_hi32), "r"(b_lo32), "r"(b_hi32), "r"(idesc), "r"(acc) : "memory");
}
template <bool A_MN, bool B_MN, int MAXK>
__device__ __forceinline__ void mma3(uint32_t d_tmem, uint32_t a_, uint32_t a_lo, uint32_t b_, uint32_t b_lo, int ksteps_,
                                     uint32_t idesc) {
  // called by a whole (converged) warp; lane 0 issues.  The broadcasts tell the compiler the operands are warp-uniform,
  // so the descriptors live in uniform registers instead of going through a per-MMA R2UR waterfall
  const uint32_t a = __shfl_sync(0xffffffffu, a_, 0), b = __shfl_sync(0xffffffffu, b_, 0);
  const int ksteps = __shfl_sync(0xffffffffu, ksteps_, 0);
  const bool leader = (threadIdx.x & 31) == 0;
  // descriptors as (low word, high word): only the low word (start address, 16-byte units) changes between MMAs
  const uint64_t a0 = A_MN ? umma_desc_mn_sw128(0, GRP) : umma_desc_sw128(0);
  const uint64_t b0 = B_MN ? umma_desc_mn_sw128(0, GRP) : umma_desc_sw128(0);
  const uint32_t a_hi32 = (uint32_t)(a0 >> 32), b_hi32 = (uint32_t)(b0 >> 32);
  const uint32_t ah = (uint32_t)a0 | ((a >> 4) & 0x3FFFu), al = (uint32_t)a0 | (((a + a_lo) >> 4) & 0x3FFFu);
  const uint32_t bh = (uint32_t)b0 | ((b >> 4) & 0x3FFFu), bl = (uint32_t)b0 | (((b + b_lo) >> 4) & 0x3FFFu);
#pragma unroll
  for (int ks = 0; ks < MAXK; ++ks) {
    if (ks < ksteps && leader) {
      const uint32_t ao = A_MN ? (uint32_t)ks * 64u : (uint32_t)((ks >> 2) * (GRP >> 4) + (ks & 3) * 2);
      const uint32_t bo = B_MN ? (uint32_t)ks * 64u : (uint32_t)((ks >> 2) * (GRP >> 4) + (ks & 3) * 2);
      umma_tf32_split(d_tmem, al + ao, a_hi32, bh + bo, b_hi32, idesc, ks ? 1u : 0u);     // small terms first
      umma_tf32_split(d_tmem, ah + ao, a_hi32, bl + bo, b_hi32, idesc, 1u);
      umma_tf32_split(d_tmem, ah + ao, a_hi32, bh + bo, b_hi32, idesc, 1u);
    }
  }
}

// four consecutive values (columns 4c..4c+3) of this thread's row i of a [64 x 64] tile: hi -> tile, lo -> tile + PT
__device__ __forceinline__ void store_chunk_hi_lo(uint32_t tile, int i, int c, float v0, float v1, float v2, float v3) {
  const uint32_t off = (uint32_t)((c >> 3) * GRP + i * 128 + (((c & 7) ^ (i & 7)) << 4));
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile + off), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(tile + PT + off), "f"(lo_of(v0)), "f"(lo_of(v1)), "f"(lo_of(v2)),
               "f"(lo_of(v3)) : "memory");
}

// same values into an image that is read column-wise (MN-major A operand): 32-byte-atom swizzle
__device__ __forceinline__ void store_chunk_mn(uint32_t hi, uint32_t lo, int i, int c, float v0, float v1, float v2, float v3) {
  const uint32_t off = (uint32_t)((c >> 3) * GRP) + mn_sw_offset(i, c & 7);
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hi + off), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(lo + off), "f"(lo_of(v0)), "f"(lo_of(v1)), "f"(lo_of(v2)),
               "f"(lo_of(v3)) : "memory");
}

// keep/scale factors (0 or 1/(1-p)) of the attention-dropout decisions (i, 4c..4c+3): index space [B, H, T, T]
__device__ __forceinline__ float4 mask4(const AttnTcP& p, const RngKey& key, uint64_t row_base, int c, float ik) {
  if (p.drop_p <= 0.f) return make_float4(1.f, 1.f, 1.f, 1.f);
  if (4 * c >= p.T) return make_float4(0.f, 0.f, 0.f, 0.f);
  if ((p.T & 3) == 0) return dropout_scale4(key, p.site, row_base + 4 * c, p.drop_p, ik);     // row_base % 4 == 0
  float4 q;
  q.x = dropout_scale(key, p.site, row_base + 4 * c, p.drop_p, ik);
  q.y = 4 * c + 1 < p.T ? dropout_scale(key, p.site, row_base + 4 * c + 1, p.drop_p, ik) : 0.f;
  q.z = 4 * c + 2 < p.T ? dropout_scale(key, p.site, row_base + 4 * c + 2, p.drop_p, ik) : 0.f;
  q.w = 4 * c + 3 < p.T ? dropout_scale(key, p.site, row_base + 4 * c + 3, p.drop_p, ik) : 0.f;
  return q;
}

// Attention-dropout keep bits, computed by all 128 threads while the tiles are still in flight: the counter-based
// stream needs nothing but indices.  Thread t covers row t & 63, columns 32*(t >> 6) .. +31 (8 Philox blocks, four
// independent ones in flight: a fully rolled loop is one 10-round dependent chain after another and took 2.8 us, a
// fully unrolled one is 1600 instructions of straight-line code).  bit j of keep[r] = element (r, j) is kept.
__device__ __forceinline__ void precompute_keep_bits(const AttnTcP& p, const RngKey& key, int b, int h, unsigned long long* keep) {
  if (p.drop_p <= 0.f) return;
  const int r = threadIdx.x & 63, half = threadIdx.x >> 6;
  const uint64_t row_base = ((uint64_t)(b * p.H + h) * p.T + r) * p.T;
  uint32_t bits = 0u;
  if (r < p.T) {
#pragma unroll 1
    for (int c0 = 0; c0 < 8; c0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 m = mask4(p, key, row_base, 8 * half + c0 + u, 1.f);
        bits |= ((m.x > 0.f ? 1u : 0u) | (m.y > 0.f ? 2u : 0u) | (m.z > 0.f ? 4u : 0u) | (m.w > 0.f ? 8u : 0u)) << (4 * (c0 + u));
      }
    }
  }
  reinterpret_cast<uint32_t*>(keep)[2 * r + half] = bits;      // little-endian halves of the 64-bit row mask
}

// Row softmax in rolled passes over 16-column chunks of the accumulator row (re-read from TMEM each pass): the kernel
// runs each phase once per CTA, so fully unrolled 64-wide code was instruction-fetch bound (~7 clocks per instruction).
// exp(x - m) = ex2((x - m) * log2 e), one MUFU per element.
struct RowStat { float mxs, inv; };      // max * (scale * log2 e), 1 / sum (0 for a padded row)
__device__ __forceinline__ float row_max(uint32_t trow, int nv) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int cc = 0; cc < 4; ++cc) {
    float v[16];
    tmem_ld16(trow + 16 * cc, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, 16 * cc + j < nv ? v[j] : -INFINITY);
  }
  return mx;
}
__device__ __forceinline__ float exp_el(float v, float sl2, float mxs, bool valid) { return valid ? ex2_approx(fmaf(v, sl2, -mxs)) : 0.f; }

// =================================================================================================
// forward: ctx[t, b, h*hd + d] = sum_j dropout(softmax(scale * Q K^T))[t, j] V[j, d]
//   shared memory: [Q | later V : hi, lo] [K | later P : hi, lo]  = 96 KB -> two CTAs per SM
// =================================================================================================
__global__ void __launch_bounds__(NTHR) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                           const __grid_constant__ CUtensorMap tmQKVm, const AttnTcP p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const uint32_t QV = base, KP = base + 2u * TILE;
  const uint32_t bar = base + 4u * TILE;
  const uint32_t bar_qk = bar, bar_v = bar + 8, bar_s = bar + 16, bar_o = bar + 24, tmem_slot = bar + 32;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  unsigned long long* keep = reinterpret_cast<unsigned long long*>(smem_raw + (bar + 64 - smem_u32(smem_raw)));   // [64]

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKVm) : "memory");
      mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const RngKey key = load_rng_key(p.drop_p > 0.f ? p.rng : nullptr);     // one read per thread (after the dependency wait: nothing global is touched before it)
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tO = tmem + 64;
  stamp(p, 0);

  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_qk, 2u * TILE);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      tma_load_5d(&tmQKV, bar_qk, QV + g * GRP, 32 * g, h, 0, b, 0);
      tma_load_5d(&tmQKV, bar_qk, KP + g * GRP, 32 * g, h, 1, b, 0);
    }
  }
  precompute_keep_bits(p, key, b, h, keep);
  stamp(p, 1);
  mbar_wait(bar_qk, 0);
  stamp(p, 2);
  lo_pass(QV, TILE, TILE);
  lo_pass(KP, TILE, TILE);
  fence_async_smem();
  __syncthreads();
  stamp(p, 3);
  if (warp == 0) {
    tc_fence_after();
    mma3<false, false, 4 * NG>(tS, QV, TILE, KP, TILE, (p.hd + 7) >> 3, umma_idesc_tf32(128, 64, false, false));
  }
  if (threadIdx.x == 0) {
    umma_commit(bar_s);
    stamp(p, 4);
    mbar_wait(bar_s, 0);                       // Q is dead: its region receives V
    stamp(p, 5);
    mbar_expect_tx(bar_v, (uint32_t)TILE);
#pragma unroll
    for (int g = 0; g < NG; ++g) tma_load_5d(&tmQKVm, bar_v, QV + g * GRP, 32 * g, h, 2, b, 0);     // MN image
  }
  // Lanes 1..31 of warp 0 must not run ahead of lane 0: divergent paths of one warp execute one at a time, so sibling
  // lanes spinning on an mbarrier (or starting their softmax) would time-slice with the MMA issue loop and lane 0 would
  // then redo the softmax alone.  Park them at a warp barrier instead.
  __syncwarp();
  if (warp < 2) {      // rows 0..63: masked softmax + dropout; P (hi, lo) replaces K in shared memory
    const int i = threadIdx.x;
    const long long len = p.lengths[b];
    const int nv = (int)(len < p.T ? (len < 0 ? 0 : len) : p.T);
    mbar_wait(bar_s, 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t trow = tS + ((uint32_t)(warp * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    const float mxs = row_max(trow, nv) * sl2;
    float sum = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      float v[16];
      tmem_ld16(trow + 16 * cc, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) sum += exp_el(v[j], sl2, mxs, 16 * cc + j < nv);
    }
    const bool drop = p.drop_p > 0.f;
    const float invk = ((i < p.T && nv > 0) ? 1.f / sum : 0.f) * (drop ? 1.f / (1.f - p.drop_p) : 1.f);
    const unsigned long long bits = drop ? keep[i] : ~0ull;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {      // P * keep / (1 - p): hi and lo images for O = P V
      float v[16];
      tmem_ld16(trow + 16 * cc, v);
      const unsigned kb = (unsigned)(bits >> (16 * cc)) & 0xFFFFu;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = 4 * c4 + q;
          const float e = exp_el(v[j], sl2, mxs, 16 * cc + j < nv);
          o[q] = (kb >> j) & 1u ? e * invk : 0.f;
        }
        store_chunk_hi_lo(KP, i, 4 * cc + c4, o[0], o[1], o[2], o[3]);
      }
    }
    fence_async_smem();
    stamp(p, 6);
  } else {      // warps 2, 3 meanwhile: remainder image of V (it lands during the softmax)
    mbar_wait(bar_v, 0);
    stamp_by(p, 7, 64);
    lo_image<6>(QV, QV + TILE, TILE / 16u, threadIdx.x - 64u, 64u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  stamp(p, 8);
  if (warp == 0) {
    tc_fence_after();
    mma3<false, true, 8>(tO, KP, PT, QV, TILE, 8, umma_idesc_tf32(128, 96, false, true));
  }
  if (threadIdx.x == 0) {
    umma_commit(bar_o);
    stamp(p, 9);
  }
  __syncwarp();
  if (warp < 2) {
    const int i = threadIdx.x;
    mbar_wait(bar_o, 0);
    __syncwarp();
    stamp(p, 10);
    tc_fence_after();
    float* dst = p.ctx + ((long long)i * p.B + b) * p.D + h * p.hd;
#pragma unroll
    for (int ch = 0; ch < NG; ++ch) {
      uint32_t v[32];
      tmem_ld32(tO + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
      if (i < p.T) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const int d = ch * 32 + 4 * q4;
          if (d < p.hd) *reinterpret_cast<uint4*>(dst + d) = make_uint4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
        }
      }
    }
  }
  stamp(p, 11);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
  stamp(p, 12);
}

// =================================================================================================
// backward (plan in the header comment)
// =================================================================================================
__global__ void __launch_bounds__(NTHR) attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                                                           const __grid_constant__ CUtensorMap tmQKVm,
                                                           const __grid_constant__ CUtensorMap tmDO,
                                                           const __grid_constant__ CUtensorMap tmDOm, const AttnTcP p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  constexpr uint32_t REG = 2u * TILE;                              // 48 KB: hi + lo image of one head slice
  const uint32_t R0 = base, R1 = base + REG, R2 = base + 2u * REG, R3 = base + 3u * REG;
  const uint32_t Pd = R0, Pd_lo = R0 + PT;                         // MN image
  const uint32_t dSk = R1;                                         // K-major image (hi, lo = +PT)
  const uint32_t dSm = R0 + 2u * PT, dSm_lo = R1 + 2u * PT;        // MN image in the two region tails
  const uint32_t bar = R3 + REG + GRP;                             // GRP bytes of pad: M = 128 over-read of the last region
  const uint32_t bar_qk = bar, bar_gv = bar + 8, bar_s = bar + 16, bar_dp = bar + 24, bar_m1 = bar + 32, bar_2 = bar + 40,
                 bar_m2 = bar + 48, bar_out = bar + 56, tmem_slot = bar + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  unsigned long long* keep = reinterpret_cast<unsigned long long*>(smem_raw + (bar + 128 - smem_u32(smem_raw)));  // [64]

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKVm) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDO) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDOm) : "memory");
      for (int k = 0; k < 8; ++k) mbar_init(bar + 8u * k, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const RngKey key = load_rng_key(p.drop_p > 0.f ? p.rng : nullptr);
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tS = tmem, tDP = tmem + 64, tDV = tmem + 128, tDQ = tmem + 224, tDK = tmem + 320;
  stamp(p, 0);

  // ---- phase 1: K-major images of Q, K, V, dO; S = Q K^T and dP = dO V^T ---------------------------
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_qk, 2u * TILE);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      tma_load_5d(&tmQKV, bar_qk, R0 + g * GRP, 32 * g, h, 0, b, 0);
      tma_load_5d(&tmQKV, bar_qk, R1 + g * GRP, 32 * g, h, 1, b, 0);
    }
    mbar_expect_tx(bar_gv, 2u * TILE);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      tma_load_5d(&tmQKV, bar_gv, R2 + g * GRP, 32 * g, h, 2, b, 0);
      tma_load_4d(&tmDO, bar_gv, R3 + g * GRP, 32 * g, h, b, 0);
    }
  }
  precompute_keep_bits(p, key, b, h, keep);
  mbar_wait(bar_qk, 0);
  stamp(p, 1);
  lo_pass(R0, TILE, TILE);
  lo_pass(R1, TILE, TILE);
  fence_async_smem();
  __syncthreads();
  if (warp == 0) {     // recompute the scores
    tc_fence_after();
    mma3<false, false, 4 * NG>(tS, R0, TILE, R1, TILE, (p.hd + 7) >> 3, umma_idesc_tf32(128, 64, false, false));
  }
  if (threadIdx.x == 0) {
    umma_commit(bar_s);
    stamp(p, 2);
  }
  __syncwarp();        // (see the forward kernel: sibling lanes must not spin while lane 0 issues MMAs)
  mbar_wait(bar_gv, 0);
  lo_pass(R2, TILE, TILE);
  lo_pass(R3, TILE, TILE);
  fence_async_smem();
  __syncthreads();
  if (warp == 0) {     // dPd[i, j] = sum_d dO[i, d] V[j, d]
    tc_fence_after();
    mma3<false, false, 4 * NG>(tDP, R3, TILE, R2, TILE, (p.hd + 7) >> 3, umma_idesc_tf32(128, 64, false, false));
  }
  if (threadIdx.x == 0) {
    umma_commit(bar_dp);
    stamp(p, 3);
    // every phase-1 image is dead once both accumulators are complete: fetch the MN images of dO and K
    mbar_wait(bar_s, 0);
    mbar_wait(bar_dp, 0);
    stamp(p, 4);
    mbar_expect_tx(bar_m1, 2u * TILE);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      tma_load_4d(&tmDOm, bar_m1, R2 + g * GRP, 32 * g, h, b, 0);
      tma_load_5d(&tmQKVm, bar_m1, R3 + g * GRP, 32 * g, h, 1, b, 0);
    }
  }
  __syncwarp();
  // ---- phase 2: softmax / dS math on rows 0..63, Pd and dS images into R0 / R1 ----------------------
  if (warp < 2) {
    const int i = threadIdx.x;
    const long long len = p.lengths[b];
    const int nv = (int)(len < p.T ? (len < 0 ? 0 : len) : p.T);
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    mbar_wait(bar_s, 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t srow = tS + lane_addr, drow = tDP + lane_addr;
    const float sl2 = p.scale * 1.4426950408889634f;
    const float mxs = row_max(srow, nv) * sl2;
    mbar_wait(bar_dp, 0);                            // Q, K, V, dO images are dead from here on
    __syncwarp();
    tc_fence_after();
    const bool drop = p.drop_p > 0.f;
    const float ik = drop ? 1.f / (1.f - p.drop_p) : 1.f;
    const unsigned long long bits = drop ? keep[i] : ~0ull;
    float sum = 0.f, dotr = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {      // sum of exponentials and rowsum(dP * P) (un-normalised) in one pass
      float v[16], g[16];
      tmem_ld16(srow + 16 * cc, v);
      tmem_ld16(drow + 16 * cc, g);
      const unsigned kb = (unsigned)(bits >> (16 * cc)) & 0xFFFFu;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e = exp_el(v[j], sl2, mxs, 16 * cc + j < nv);
        sum += e;
        dotr += (kb >> j) & 1u ? g[j] * e : 0.f;
      }
    }
    const float inv = (i < p.T && nv > 0) ? 1.f / sum : 0.f;
    const float dot = dotr * ik * inv;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      float v[16], g[16];
      tmem_ld16(srow + 16 * cc, v);
      tmem_ld16(drow + 16 * cc, g);
      const unsigned kb = (unsigned)(bits >> (16 * cc)) & 0xFFFFu;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        float pd[4], ds[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = 4 * c4 + q;
          const float pr = exp_el(v[j], sl2, mxs, 16 * cc + j < nv) * inv;      // probability
          const float m = (kb >> j) & 1u ? ik : 0.f;
          pd[q] = pr * m;                                                        // Pd = P * mask (read column-wise by dV)
          ds[q] = pr * (g[j] * m - dot) * p.scale;                               // scale * dS = scale * P * (dP - rowsum(dP * P))
        }
        const int c = 4 * cc + c4;
        store_chunk_mn(Pd, Pd_lo, i, c, pd[0], pd[1], pd[2], pd[3]);
        store_chunk_hi_lo(dSk, i, c, ds[0], ds[1], ds[2], ds[3]);                // row-wise image (dQ = dS K)
        store_chunk_mn(dSm, dSm_lo, i, c, ds[0], ds[1], ds[2], ds[3]);           // column-wise image (dK = dS^T Q)
      }
    }
    fence_async_smem();
    stamp(p, 5);
  } else {      // warps 2, 3 have no accumulator rows: they derive the remainder images of the MN tiles meanwhile
    mbar_wait(bar_m1, 0);
    stamp_by(p, 6, 64);
    lo_image<6>(R2, R2 + TILE, TILE / 16u, threadIdx.x - 64u, 64u);
    lo_image<6>(R3, R3 + TILE, TILE / 16u, threadIdx.x - 64u, 64u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    // dV[j, d] = sum_i Pd[i, j] dO[i, d]        A = Pd read column-wise (MN-major), B = dO (MN-major)
    mma3<true, true, 8>(tDV, Pd, PT, R2, TILE, 8, umma_idesc_tf32(128, 96, true, true));
    // dQ[i, d] = sum_j (scale dS)[i, j] K[j, d]
    mma3<false, true, 8>(tDQ, dSk, PT, R3, TILE, 8, umma_idesc_tf32(128, 96, false, true));
  }
  if (threadIdx.x == 0) {
    umma_commit(bar_2);
    stamp(p, 7);
    mbar_wait(bar_2, 0);                 // the dO image is dead: its region receives Q (MN image)
    mbar_expect_tx(bar_m2, (uint32_t)TILE);
#pragma unroll
    for (int g = 0; g < NG; ++g) tma_load_5d(&tmQKVm, bar_m2, R2 + g * GRP, 32 * g, h, 0, b, 0);
    stamp(p, 8);
  }
  __syncwarp();
  auto store_out = [&](uint32_t t0, int which) {     // accumulator rows 0..63 -> d_qkv[t, b, which*D + h*hd + d]
    const int i = threadIdx.x;
    float* dst = p.dqkv + ((long long)i * p.B + b) * 3 * p.D + h * p.hd + which * p.D;
#pragma unroll
    for (int ch = 0; ch < NG; ++ch) {
      uint32_t v[32];
      tmem_ld32(t0 + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
      if (i < p.T) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const int d = ch * 32 + 4 * q4;
          if (d < p.hd) *reinterpret_cast<uint4*>(dst + d) = make_uint4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
        }
      }
    }
  };
  if (warp < 2) {                        // dQ and dV leave while warps 2, 3 prepare and issue the last product
    mbar_wait(bar_2, 0);
    __syncwarp();
    tc_fence_after();
    store_out(tDQ, 0);
    store_out(tDV, 2);
    stamp(p, 9);
  } else {
    // ---- phase 3: dK[j, d] = sum_i (scale dS)[i, j] Q[i, d] -------------------------------------------
    mbar_wait(bar_m2, 0);
    stamp_by(p, 10, 64);
    lo_image<6>(R2, R2 + TILE, TILE / 16u, threadIdx.x - 64u, 64u);
    fence_async_smem();
    asm volatile("bar.sync 1, 64;" ::: "memory");      // warps 2 and 3
    if (warp == 2) {
      tc_fence_after();
      mma3<true, true, 8>(tDK, dSm, dSm_lo - dSm, R2, TILE, 8, umma_idesc_tf32(128, 96, true, true));
      if (lane == 0) umma_commit(bar_out);
      stamp_by(p, 11, 64);
      __syncwarp();
    }
  }
  if (warp < 2) {
    mbar_wait(bar_out, 0);
    __syncwarp();
    tc_fence_after();
    store_out(tDK, 1);
  }
  stamp(p, 12);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

constexpr int FWD_SMEM = 1024 + 4 * TILE + 64 + 512;
constexpr int BWD_SMEM = 1024 + 8 * TILE + GRP + 128 + 512;

// qkv viewed as [T, B, 3, H, hd]: box = 32 columns x 64 timestamps of one (sample, q/k/v, head)
int encode_qkv(CUtensorMap* m, const float* qkv, int B, int H, int T, int hd, CUtensorMapSwizzle sw) {
  const cuuint64_t D = (cuuint64_t)H * hd;
  cuuint64_t dims[5] = {(cuuint64_t)hd, (cuuint64_t)H, 3, (cuuint64_t)B, (cuuint64_t)T};
  cuuint64_t str[4] = {(cuuint64_t)hd * 4, D * 4, 3 * D * 4, (cuuint64_t)B * 3 * D * 4};
  cuuint32_t box[5] = {32, 1, 1, 1, TR};
  return encode(m, qkv, 5, dims, str, box, sw, "attention qkv");
}
int encode_ctx(CUtensorMap* m, const float* x, int B, int H, int T, int hd, CUtensorMapSwizzle sw) {
  const cuuint64_t D = (cuuint64_t)H * hd;
  cuuint64_t dims[4] = {(cuuint64_t)hd, (cuuint64_t)H, (cuuint64_t)B, (cuuint64_t)T};
  cuuint64_t str[3] = {(cuuint64_t)hd * 4, D * 4, (cuuint64_t)B * D * 4};
  cuuint32_t box[4] = {32, 1, 1, TR};
  return encode(m, x, 4, dims, str, box, sw, "attention d(ctx)");
}

}  // namespace

static unsigned long long* g_attn_dbg = nullptr;
void attn_tc_set_debug(unsigned long long* buf) { g_attn_dbg = buf; }

bool attn_tc_supported(int T, int hd) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("RD_ATTN_TC"); env = (e && e[0] == '0') ? 0 : 1; }
  return env == 1 && T <= TR && hd <= 32 * NG && hd % 4 == 0 && hd >= 4;
}

int attn_tc_fwd(const float* qkv, const int64_t* lengths, int B, int H, int T, int hd, float drop_p, const uint64_t* rng,
                uint32_t site, float* ctx, cudaStream_t st) {
  if (!attn_tc_supported(T, hd) || ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(ctx)) & 15)) {
    set_error("attn_tc_fwd: unsupported shape / alignment (T=%d hd=%d)", T, hd);
    return -2;
  }
  AttnTcP p{};
  p.ctx = ctx; p.lengths = lengths; p.B = B; p.H = H; p.T = T; p.hd = hd; p.D = H * hd;
  p.scale = 1.f / sqrtf((float)hd); p.drop_p = drop_p; p.rng = rng; p.site = site; p.dbg = g_attn_dbg;
  CUtensorMap tm, tmm;
  RD_TRY(encode_qkv(&tm, qkv, B, H, T, hd, CU_TENSOR_MAP_SWIZZLE_128B));
  RD_TRY(encode_qkv(&tmm, qkv, B, H, T, hd, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  RD_TRY(ensure_max_smem((const void*)attn_tc_fwd_kernel, FWD_SMEM));
  launch_pdl(attn_tc_fwd_kernel, dim3(B * H), dim3(NTHR), FWD_SMEM, st, tm, tmm, p);
  RD_CHECK_LAUNCH("attn_tc_fwd_kernel");
  return 0;
}

int attn_tc_bwd(const float* qkv, const float* dctx, const int64_t* lengths, int B, int H, int T, int hd, float drop_p,
                const uint64_t* rng, uint32_t site, float* dqkv, cudaStream_t st) {
  if (!attn_tc_supported(T, hd) ||
      ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(dctx) | reinterpret_cast<uintptr_t>(dqkv)) & 15)) {
    set_error("attn_tc_bwd: unsupported shape / alignment (T=%d hd=%d)", T, hd);
    return -2;
  }
  AttnTcP p{};
  p.dqkv = dqkv; p.lengths = lengths; p.B = B; p.H = H; p.T = T; p.hd = hd; p.D = H * hd;
  p.scale = 1.f / sqrtf((float)hd); p.drop_p = drop_p; p.rng = rng; p.site = site; p.dbg = g_attn_dbg;
  CUtensorMap tq, tqm, tg, tgm;
  RD_TRY(encode_qkv(&tq, qkv, B, H, T, hd, CU_TENSOR_MAP_SWIZZLE_128B));
  RD_TRY(encode_qkv(&tqm, qkv, B, H, T, hd, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  RD_TRY(encode_ctx(&tg, dctx, B, H, T, hd, CU_TENSOR_MAP_SWIZZLE_128B));
  RD_TRY(encode_ctx(&tgm, dctx, B, H, T, hd, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  RD_TRY(ensure_max_smem((const void*)attn_tc_bwd_kernel, BWD_SMEM));
  launch_pdl(attn_tc_bwd_kernel, dim3(B * H), dim3(NTHR), BWD_SMEM, st, tq, tqm, tg, tgm, p);
  RD_CHECK_LAUNCH("attn_tc_bwd_kernel");
  return 0;
}

}  // namespace rd
