// fp32-accurate GEMM on the 5th-gen tensor cores for the temporal self-attention encoder: the dense
// Wq/Wk/Wv/Wo and feed-forward projections (code/models_rd.py:232-237,358 -> nn.TransformerEncoder)
// and their input gradients.  Single-pass TF32 is borderline for these layers (SURVEY.md section 7:
// ~1e-3 on the logits), so every product is error-compensated: the MMA reads the top 19 bits of an
// fp32 operand (= its "hi" part) and we feed the exact remainders as two more MMAs into the same TMEM
// accumulator.  Warp roles (10 warps, one persistent CTA per SM):
//   warp 0      TMA producer for the weight tile and its precomputed remainder (L2 resident)
//   warp 1      one thread issuing 3 x tcgen05.mma.kind::tf32 per 8-wide k-step, TMEM accumulators
//               double buffered so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2-5   epilogue: tcgen05.ld -> bias / relu / gate / dropout / residual -> swizzled smem ->
//               TMA store (coalesced 128-byte lines)
//   warps 6-9   stagers: the activation tile arrives by TMA like the weight tile; these warps move it from shared memory
//               into TENSOR memory (tcgen05.st, thread = row), split into hi (raw) and lo = x - trunc19(x).  The MMAs then
//               take A from tensor memory: with both operands in shared memory the three error-compensation passes read
//               the 128-row A slice three times per k-step and the kernel was shared-memory-bandwidth bound (156 KB of
//               shared-memory traffic per k-block at BN = 96, 0.56 us; the tensor pipe idle 97 % of the time)
#include <stdlib.h>

#include "rd_tc_common.cuh"
#include "rd_tc_gemm.cuh"

namespace rd {
using namespace tc;
namespace {

constexpr int BM = 128, BK = 32, MAX_STAGES = 4, NTHREADS = 320;
constexpr int A_TILE = BM * BK * 4;        // 16 KB (hi) + 16 KB (lo)
constexpr int STG_BYTES = 4096;
constexpr int MAX_BN = 160;
// tensor memory: two accumulators of ACC_COLS columns, then A_SLOTS activation slots of 64 columns (hi | lo, 32 each)
constexpr uint32_t ACC_COLS = 160, A_COL0 = 2 * ACC_COLS;
constexpr int A_SLOTS = 3;

struct P {
  const float* A; long long lda;
  long long M; int N, K, BN, n_tiles, m_tiles, k_blocks, nstages;
  const float* bias; int relu;
  const float* gate; long long gate_ld; float gate_scale;
  float drop_p; const uint64_t* rng; uint32_t drop_site;
  uint32_t* drop_mask; int drop_mask_ld;
  const float* resid; long long resid_ld;
  unsigned long long* dbg;     // optional %globaltimer phase stamps [CTA][8] (rd_debug_gemm_timing)
};

__device__ __forceinline__ void gstamp(const P& p, int slot) {
  if (p.dbg) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.dbg[(size_t)blockIdx.x * 8 + slot] = t;
  }
}

// epilogue features are compile-time: the epilogue is on the critical path of these small GEMMs
template <bool RELU, bool GATE, bool DROP, bool RESID>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmC, const P p) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  if (threadIdx.x == 0) gstamp(p, 0);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b_tile = (uint32_t)p.BN * 128u;
  const uint32_t stage_bytes = (uint32_t)A_TILE + 2u * b_tile;      // A (raw) | B hi | B lo
  const uint32_t stg_base = base + (uint32_t)p.nstages * stage_bytes;
  const uint32_t bias_base = stg_base + 8 * STG_BYTES;
  const uint32_t bar_base = bias_base + 2 * 256 * 4;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);
  auto aready_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 5 + a); };    // A slot staged in tensor memory
  auto aempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 5 + A_SLOTS + a); };   // MMAs have read it
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bias_base - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int a = 0; a < A_SLOTS; ++a) { mbar_init(aready_bar(a), 4); mbar_init(aempty_bar(a), 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();          // everything above is private to this CTA; the previous kernel's output is first touched below
  if (threadIdx.x == 0) gstamp(p, 1);
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ===== TMA producer: activation tile, weight tile (hi = the raw weight) and the weight remainder ==============
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_t = tile % p.n_tiles, m_t = tile / p.n_tiles;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), (uint32_t)A_TILE + 2u * b_tile);
          tma_load_2d(&tmA, full_bar(stage), base + (uint32_t)stage * stage_bytes, kb * BK, m_t * BM);
          const uint32_t sb = base + (uint32_t)stage * stage_bytes + (uint32_t)A_TILE;
          tma_load_2d(&tmB, full_bar(stage), sb, kb * BK, n_t * p.BN);
          tma_load_2d(&tmBlo, full_bar(stage), sb + b_tile, kb * BK, n_t * p.BN);     // (deriving it on chip like A_lo was slower:
          // the remainder pass, not L2, paces the k-loop -- 0.73 vs 0.56 us per k-block)
          if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ================================================================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0; int slot = 0; uint32_t slot_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * ACC_COLS;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);           // weight tiles (the stagers waited for the same barrier)
          mbar_wait(aready_bar(slot), slot_phase);     // activation tile staged in tensor memory
          if (kb == 0) gstamp(p, 2);
          tc_fence_after();
          const uint32_t sb = base + (uint32_t)stage * stage_bytes + (uint32_t)A_TILE;
          const uint64_t b_hi = umma_desc_sw128(sb), b_lo = umma_desc_sw128(sb + b_tile);
          const uint32_t a_hi = tmem_base + A_COL0 + (uint32_t)slot * 64u, a_lo = a_hi + 32u;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t o = (uint64_t)(kk * 2);
            const uint32_t ao = (uint32_t)(kk * 8);                                  // 8 columns = 8 tf32 of K
            umma_tf32_ts(d_tmem, a_lo + ao, b_hi + o, idesc, (kb | kk) ? 1u : 0u);   // small terms first
            umma_tf32_ts(d_tmem, a_hi + ao, b_lo + o, idesc, 1u);
            umma_tf32_ts(d_tmem, a_hi + ao, b_hi + o, idesc, 1u);
          }
          umma_commit(empty_bar(stage));
          umma_commit(aempty_bar(slot));
          if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
          if (++slot == A_SLOTS) { slot = 0; slot_phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        gstamp(p, 3);
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < 6) {
    // ===== epilogue ====================================================================================
    const int q = warp & 3;
    const int et = threadIdx.x - 64;
    const uint32_t my_stg = stg_base + (uint32_t)(warp - 2) * 2u * STG_BYTES;
    int acc = 0; uint32_t acc_phase = 0; int buf = 0;
    const int n_chunks = (p.BN + 31) / 32;
    const float ik = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const RngKey key = load_rng_key(DROP ? p.rng : nullptr);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_t = tile / p.n_tiles, n_t = tile - m_t * p.n_tiles;
      const int col0 = n_t * p.BN;
      for (int c = et; c < 256; c += 128) bias_s[acc * 256 + c] = (p.bias && c < p.BN && col0 + c < p.N) ? __ldg(p.bias + col0 + c) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int row0 = m_t * BM + q * 32;
      const long long row = (long long)row0 + lane;
      const bool row_ok = row < p.M;
      // gate / residual operands of chunk ch are fetched one chunk AHEAD (chunk 0 before the accumulator is even
      // complete): their L2 round trips used to sit, one per chunk, on the critical path of the epilogue
      float4 pg[8], pr[8];
      auto prefetch = [&](int ch) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const int c = col0 + ch * 32 + 4 * j4;
          const bool ok = row_ok && c < p.N && ch * 32 + 4 * j4 < p.BN;
          if (GATE) pg[j4] = ok ? __ldg(reinterpret_cast<const float4*>(p.gate + row * p.gate_ld + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (RESID) pr[j4] = ok ? __ldg(reinterpret_cast<const float4*>(p.resid + row * p.resid_ld + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      if (GATE || RESID) prefetch(0);
      mbar_wait(tfull_bar(acc), acc_phase);
      if (threadIdx.x == 64) gstamp(p, 4);
      tc_fence_after();
      for (int ch = 0; ch < n_chunks; ++ch) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * ACC_COLS + (uint32_t)(ch * 32), v);
        float4 cg[8], cr[8];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) { if (GATE) cg[j4] = pg[j4]; if (RESID) cr[j4] = pr[j4]; }
        if ((GATE || RESID) && ch + 1 < n_chunks) prefetch(ch + 1);
        const float* bs = bias_s + acc * 256 + ch * 32;
        const int c0 = col0 + ch * 32;
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        const uint32_t stg = my_stg + (uint32_t)buf * STG_BYTES;
        uint32_t keep_word = 0u;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float o[4];
          const int c = c0 + 4 * j4;
          const bool ok = row_ok && c < p.N;     // N % 4 == 0: a float4 is all in or all out
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x = __uint_as_float(v[4 * j4 + e]) + bs[4 * j4 + e];
            o[e] = RELU ? fmaxf(x, 0.f) : x;
          }
          if (GATE && ok) {
            const float4 g = cg[j4];
            o[0] = g.x > 0.f ? o[0] * p.gate_scale : 0.f; o[1] = g.y > 0.f ? o[1] * p.gate_scale : 0.f;
            o[2] = g.z > 0.f ? o[2] * p.gate_scale : 0.f; o[3] = g.w > 0.f ? o[3] * p.gate_scale : 0.f;
          }
          if (DROP && ok) {   // row*N + c is a multiple of 4: one Philox block for the four columns
            const float4 m = dropout_scale4(key, p.drop_site, (uint64_t)row * (uint64_t)p.N + (uint64_t)c, p.drop_p, ik);
            o[0] *= m.x; o[1] *= m.y; o[2] *= m.z; o[3] *= m.w;
            keep_word |= ((m.x > 0.f ? 1u : 0u) | (m.y > 0.f ? 2u : 0u) | (m.z > 0.f ? 4u : 0u) | (m.w > 0.f ? 8u : 0u)) << (4 * j4);
          }
          if (RESID && ok) {
            const float4 r = cr[j4];
            o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
          }
          const uint32_t off = (uint32_t)(lane * 128 + ((j4 ^ (lane & 7)) << 4));
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + off), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
        }
        // the backward of the consumer (LayerNorm) reads these bits instead of regenerating the Philox stream
        if (DROP && p.drop_mask && row_ok && c0 < p.N) p.drop_mask[row * p.drop_mask_ld + (c0 >> 5)] = keep_word;
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmC, stg, c0, row0);
          bulk_commit();
        }
        buf ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1; if (acc == 0) acc_phase ^= 1u;
    }
    if (threadIdx.x == 64) gstamp(p, 5);
    if (lane == 0) bulk_wait_read<0>();
    if (threadIdx.x == 64) gstamp(p, 6);
  } else {
    // ===== stagers: activation tile shared memory -> tensor memory, split into hi (raw) and lo = x - trunc19(x) ==========
    // thread = tile row (its TMEM lane); the row's 32 values of the k-block are 8 swizzled 16-byte pieces
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int stage = 0; uint32_t phase = 0; int slot = 0; uint32_t slot_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        const uint32_t sa = base + (uint32_t)stage * stage_bytes + (uint32_t)row * 128u;
        uint32_t x[32], l[32];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x[4 * c]), "=r"(x[4 * c + 1]), "=r"(x[4 * c + 2]), "=r"(x[4 * c + 3])
                       : "r"(sa + (uint32_t)(((c ^ (row & 7)) << 4))));
#pragma unroll
        for (int e = 0; e < 32; ++e) l[e] = __float_as_uint(__uint_as_float(x[e]) - __uint_as_float(x[e] & 0xFFFFE000u));
        mbar_wait(aempty_bar(slot), slot_phase ^ 1u);      // the MMAs of three k-blocks ago have read this slot
        tc_fence_after();
        const uint32_t ta = tmem_base + lane_addr + A_COL0 + (uint32_t)slot * 64u;
        tmem_st32(ta, x);
        tmem_st32(ta + 32u, l);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(aready_bar(slot));
        if (++stage == p.nstages) { stage = 0; phase ^= 1u; }
        if (++slot == A_SLOTS) { slot = 0; slot_phase ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    if (lane == 0) gstamp(p, 7);
  }
}

// =================================================================================================
// weight-gradient kernel ("TN"), GROUPED: one launch serves up to WG_MAX independent problems
//   D_k[m, n] = sum_r A_k[r, m] * B_k[r, n]   over this CTA's row range of problem k
// A training step has ten of these (8 encoder weights + 2 lin_value); launched one by one each was a
// ~20 us latency chain (2 k-blocks per CTA).  Grouped, every CTA owns >= 8 k-blocks, the loader warps
// prefetch the next k-block's global loads before they transpose/store the current one, and the
// launch + prologue latency is paid once.
// =================================================================================================
// Both operands are row-major activations [rows, cols] and the contraction runs over ROWS, i.e. they are
// "MN-major" from the tensor core's point of view.  tcgen05 takes MN-major TF32 operands directly (instruction
// descriptor bits 15/16) in exactly one shared-memory layout, the 128B swizzle with 32-byte atoms (rd_tc_common.cuh),
// which TMA produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  So the tiles go global -> shared memory exactly as
// they lie in memory: box {32 columns, 32 rows}; one 128-byte line = 32 consecutive columns of one row; the next 32
// columns are the next box, LBO = 4096 B apart.  No register transposes; four warps only derive the
// error-compensation remainders lo = x - trunc19(x) (same addresses, so the swizzle never has to be undone) and
// plant the "ones" column that makes the bias gradient fall out of the same MMAs.
struct WP {
  float* partial;                    // [nsplit][Mpad][Nld]
  long long rows; int M, N, BN, n_tiles, m_tiles, nsplit, rows_per_split, Mpad, Nld;
  int item0;                         // first work item of this problem inside the grouped list
};
struct WGroup { CUtensorMap tmA[WG_MAX]; CUtensorMap tmB[WG_MAX]; WP it[WG_MAX]; int n, total_items, nstages, stage_bytes; };

// PERSISTENT: one CTA per SM walks the work items (problem, m tile, n tile, row split) round-robin.  The accumulator is
// double-buffered in tensor memory (2 x 256 columns), so the epilogue of item i (TMEM -> partial slab, 80 KB of stores)
// overlaps the MMAs of item i+1, and the TMA ring runs ahead across item boundaries.  One CTA per item (the previous
// design) paid prologue + pipeline fill + epilogue serially per item and ran 3.1 waves as 4.
constexpr int W_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, warps 6-9 remainder pass
constexpr int MN_BOX = 32 * 32 * 4;  // one {32 col, 32 row} fp32 box = 4096 bytes

struct WItem { int pi, n_t, m_t, split, k_blocks; long long r_begin, r_end; };
__device__ __forceinline__ WItem wgrad_item(const WGroup& g, int w) {
  WItem it;
  it.pi = 0;
#pragma unroll 1
  for (int k = 1; k < g.n; ++k) if (w >= g.it[k].item0) it.pi = k;
  const WP& p = g.it[it.pi];
  const int item = w - p.item0;
  it.n_t = item % p.n_tiles; it.m_t = (item / p.n_tiles) % p.m_tiles; it.split = item / (p.n_tiles * p.m_tiles);
  it.r_begin = (long long)it.split * p.rows_per_split;
  it.r_end = min(p.rows, it.r_begin + p.rows_per_split);
  it.k_blocks = (int)((it.r_end - it.r_begin + BK - 1) / BK);
  return it;
}

__global__ void __launch_bounds__(W_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ WGroup g) {
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage_bytes = (uint32_t)g.stage_bytes;        // sized for the widest problem of the group
  const int nstages = g.nstages;
  const uint32_t bar_base = base + (uint32_t)nstages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                    // TMA bytes landed
  auto ready_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };    // remainders written
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };// MMAs have read the stage
  auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * MAX_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    for (int k = 0; k < g.n; ++k) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmA[k]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmB[k]) : "memory");
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(ready_bar(s), 4); mbar_init(empty_bar(s), 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
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
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer: raw row-major tiles, up to `nstages` k-blocks in flight ======================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < g.total_items; w += gridDim.x) {
        const WItem it = wgrad_item(g, w);
        const WP& p = g.it[it.pi];
        const int b_groups = p.BN >> 5;                       // 32-column groups of the B tile
        for (int kb = 0; kb < it.k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), (uint32_t)(4 + b_groups) * MN_BOX);
          const uint32_t sa = base + (uint32_t)stage * stage_bytes;
          const int r0 = (int)(it.r_begin + (long long)kb * BK);
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) tma_load_2d(&g.tmA[it.pi], full_bar(stage), sa + gq * MN_BOX, it.m_t * BM + gq * 32, r0);
          for (int gq = 0; gq < b_groups; ++gq)
            tma_load_2d(&g.tmB[it.pi], full_bar(stage), sa + 2u * A_TILE + gq * MN_BOX, it.n_t * p.BN + gq * 32, r0);
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ================================================================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < g.total_items; w += gridDim.x) {
        const WItem it = wgrad_item(g, w);
        const WP& p = g.it[it.pi];
        const uint32_t b_tile = (uint32_t)p.BN * 128u;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |     // MN-major A and B
                               ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        for (int kb = 0; kb < it.k_blocks; ++kb) {
          mbar_wait(ready_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)stage * stage_bytes;
          const uint64_t a_hi = umma_desc_mn_sw128(sa, MN_BOX), a_lo = umma_desc_mn_sw128(sa + A_TILE, MN_BOX);
          const uint64_t b_hi = umma_desc_mn_sw128(sa + 2u * A_TILE, MN_BOX), b_lo = umma_desc_mn_sw128(sa + 2u * A_TILE + b_tile, MN_BOX);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t o = (uint64_t)(kk * 64);            // next 8-row K group: +1024 bytes
            umma_tf32(d_tmem, a_lo + o, b_hi + o, idesc, (kb | kk) ? 1u : 0u);
            umma_tf32(d_tmem, a_hi + o, b_lo + o, idesc, 1u);
            umma_tf32(d_tmem, a_hi + o, b_hi + o, idesc, 1u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp < 6) {
    // ===== epilogue: accumulator tile -> partial slab (each lane owns one row: 128-byte runs) ===========
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < g.total_items; w += gridDim.x) {
      const WItem it = wgrad_item(g, w);
      const WP& p = g.it[it.pi];
      const int n_chunks = (p.BN + 31) / 32;
      float* prow = p.partial + ((long long)it.split * p.Mpad + it.m_t * BM + q * 32 + lane) * p.Nld + it.n_t * p.BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      for (int ch = 0; ch < n_chunks; ++ch) {
        uint32_t v[32];
        if (it.k_blocks > 0) {
          tmem_ld32(tmem_base + (uint32_t)acc * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), v);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const int c = ch * 32 + 4 * j4;
          if (c < p.BN)
            *reinterpret_cast<uint4*>(prow + c) = make_uint4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1; if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ===== warps 6-9: remainder pass per k-block (+ the implicit ones column of the bias gradient) ========
    const int lt = threadIdx.x - 192;            // 0..127
    int stage = 0; uint32_t phase = 0;
    for (int w = blockIdx.x; w < g.total_items; w += gridDim.x) {
      const WItem it = wgrad_item(g, w);
      const WP& p = g.it[it.pi];
      const uint32_t b_tile = (uint32_t)p.BN * 128u;
      const int ones_col = p.N - it.n_t * p.BN;    // tile-local column of the implicit ones column (bias gradient)
      const bool has_ones = ones_col >= 0 && ones_col < p.BN;
      const uint32_t a_vec = A_TILE / 16, b_vec = b_tile / 16;
      for (int kb = 0; kb < it.k_blocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        const uint32_t sa = base + (uint32_t)stage * stage_bytes;
        if (has_ones && lt < BK) {             // X[r, N] := 1 for the valid rows of this k-block (OOB columns arrived as 0)
          const long long r = it.r_begin + (long long)kb * BK + lt;
          const uint32_t off = (uint32_t)((ones_col >> 5) * MN_BOX) + mn_sw_offset(lt, (ones_col & 31) >> 2) + (uint32_t)(ones_col & 3) * 4u;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sa + 2u * A_TILE + off), "f"(r < it.r_end ? 1.f : 0.f) : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        lo_image<6>(sa, sa + A_TILE, a_vec, (uint32_t)lt, 128u);                              // dY tile
        lo_image<6>(sa + 2u * A_TILE, sa + 2u * A_TILE + b_tile, b_vec, (uint32_t)lt, 128u);  // X tile
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(ready_bar(stage));
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// dW[m, n] = sum_s partial[s][m][n] (n < N), db[m] = sum_s partial[s][m][N]; fixed order -> deterministic.
// One launch for every problem of a group.
// kind 0: a weight gradient (above).  kind 1: plain column sums out[c] = sum_s partial[s*stride + c], c < N
// (the LayerNorm dgamma / dbeta partial rows ride along in the same launch).
struct RItem { const float* partial; float* dW; float* db; int nsplit, M, N, Mpad, Nld, kind; long long stride; long long start; };
constexpr int RG_MAX = WG_MAX + CS_MAX;
struct RGroup { RItem it[RG_MAX]; long long total; int n; };
__global__ void wgrad_reduce_kernel(const __grid_constant__ RGroup g) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.total) return;
  int k = 0;
#pragma unroll 1
  for (int j = 1; j < g.n; ++j) if (i >= g.it[j].start) k = j;
  const RItem& r = g.it[k];
  const long long e = i - r.start;
  if (r.kind == 1) {      // one warp per output column (item starts are multiples of 32): lanes stride over the chunks
    const long long col = e >> 5;
    const int lane = (int)(e & 31);
    const float* src = r.partial + col;
    float s = 0.f;
#pragma unroll 4
    for (int sp = lane; sp < r.nsplit; sp += 32) s += __ldg(src + sp * r.stride);
    s = warp_sum(s);        // fixed butterfly order: deterministic
    if (lane == 0) r.dW[col] = s;
    return;
  }
  // kind 0: one thread per four consecutive columns of a row (N % 4 == 0, slabs 16-byte aligned), plus one per row for
  // the bias column N
  const int q_per_row = (r.N >> 2) + 1;
  const int m = (int)(e / q_per_row), q = (int)(e - (long long)m * q_per_row);
  if (m >= r.M) return;       // padding between this item and the next (warp-aligned) one
  const long long stride = (long long)r.Mpad * r.Nld;
  if (q == (r.N >> 2)) {
    const float* src = r.partial + (long long)m * r.Nld + r.N;
    float s = 0.f;
#pragma unroll 4
    for (int sp = 0; sp < r.nsplit; ++sp) s += __ldg(src + sp * stride);
    if (r.db) r.db[m] = s;
    return;
  }
  const float* src = r.partial + (long long)m * r.Nld + 4 * q;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sp0 = 0; sp0 < r.nsplit; sp0 += 8) {      // eight slabs in flight per round trip, summed in slab order
    float4 t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      t[u] = sp0 + u < r.nsplit ? __ldg(reinterpret_cast<const float4*>(src + (sp0 + u) * stride)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) { s.x += t[u].x; s.y += t[u].y; s.z += t[u].z; s.w += t[u].w; }
  }
  float* dst = r.dW + (long long)m * r.N + 4 * q;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    *reinterpret_cast<float4*>(dst) = s;
  } else {
    dst[0] = s.x; dst[1] = s.y; dst[2] = s.z; dst[3] = s.w;
  }
}

struct WPlan { int BN, n_tiles, m_tiles, nsplit, rows_per_split, Mpad, Nld; };
// `rows_target`: contraction rows per work item.  512 sizes the partial buffers (the CAPACITY in row splits); the grouped
// launch may raise it so that the item count of the whole group fills whole rounds of the persistent grid.
WPlan wgrad_plan(int M, int N, long long rows, int rows_target = 512) {
  WPlan w;
  w.m_tiles = (int)ceil_div(M, BM);
  w.Mpad = w.m_tiles * BM;
  const int Ncols = N + 1;                                  // + the ones column
  w.n_tiles = (int)ceil_div(Ncols, MAX_BN);
  w.BN = (int)round_up(ceil_div(Ncols, w.n_tiles), 32);     // whole 32-column TMA boxes
  w.Nld = (int)round_up(w.n_tiles * w.BN, 4);
  // row splits: ~512 rows (16 k-blocks) per item so the pipeline amortises its fill, but never more than
  // ~2 rounds of items per problem (bounds the partial buffer for the big configurations)
  const int mn = w.m_tiles * w.n_tiles;
  long long ns = ceil_div(rows, rows_target);
  const long long cap = (2 * num_sms()) / mn > 1 ? (2 * num_sms()) / mn : 1;
  if (ns > cap) ns = cap;
  if (ns < 1) ns = 1;
  w.rows_per_split = (int)round_up(ceil_div(rows, ns), BK);
  w.nsplit = (int)ceil_div(rows, w.rows_per_split);
  return w;
}

struct SplitItems { WeightSplit it[16]; long long start[17]; int n; StepPrologue pro; };

__global__ void split_weights_kernel(const __grid_constant__ SplitItems s) {
  pdl_launch_dependents();
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {     // step prologue riding along: dropout counter capture (+ advance) and ticket reset
    if (s.pro.rng_state) {
      s.pro.rng_captured[0] = s.pro.rng_state[0];
      s.pro.rng_captured[1] = s.pro.rng_state[1];
      if (s.pro.advance) s.pro.rng_state[1] = s.pro.rng_state[1] + 1;
    }
    if (s.pro.zero_counter) *s.pro.zero_counter = 0u;
  }
  if (i >= s.start[s.n]) return;
  int t = 0;
  while (t + 1 < s.n && i >= s.start[t + 1]) ++t;
  const WeightSplit w = s.it[t];
  long long e = i - s.start[t];
  int r = (int)(e / w.cols), c = (int)(e - (long long)r * w.cols);
  float v = w.w[e];
  float lo = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  if (w.lo) w.lo[e] = lo;
  if (w.t) { w.t[(long long)c * w.rows + r] = v; w.t_lo[(long long)c * w.rows + r] = lo; }
  if (w.rn || w.rn_t) {
    const float rn = rn_tf32(v);
    if (w.rn) w.rn[e] = rn;
    if (w.rn_t) w.rn_t[(long long)c * w.rows + r] = rn;
  }
}

void plan(long long M, int N, int* BN, int* n_tiles) {
  const long long m_tiles = ceil_div(M, BM);
  int nt = (int)ceil_div(N, MAX_BN);
  while (m_tiles * nt < 120 && round_up(ceil_div(N, nt + 1), 32) >= 64) ++nt;   // spread over the SMs
  *n_tiles = nt;
  *BN = nt == 1 ? (int)round_up(N, 16) : (int)round_up(ceil_div(N, nt), 32);
}


int ensure_attr(const void* fn, int) { return ensure_max_smem(fn, SMEM_LIMIT); }

}  // namespace

static unsigned long long* g_gemm_dbg = nullptr;
void tc_gemm_set_debug(unsigned long long* buf) { g_gemm_dbg = buf; }

bool tc_gemm_supported(const TcGemmArgs& a) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("RD_TC_GEMM"); env = (e && e[0] == '0') ? 0 : 1; }
  if (env != 1) return false;
  if (a.K % 4 || a.N % 4 || a.lda % 4 || a.K < 8 || a.N < 16 || a.M < 1 || a.M > 0x7fffffffLL) return false;
  if ((a.gate && a.gate_ld % 4) || (a.resid && a.resid_ld % 4)) return false;
  uintptr_t bits = reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.B) | reinterpret_cast<uintptr_t>(a.B_lo) |
                   reinterpret_cast<uintptr_t>(a.C) | reinterpret_cast<uintptr_t>(a.gate) | reinterpret_cast<uintptr_t>(a.resid);
  const int id = (a.relu ? 8 : 0) | (a.gate ? 4 : 0) | (a.drop_p > 0.f ? 2 : 0) | (a.resid ? 1 : 0);
  const bool combo = id == 0 || id == 1 || id == 2 || id == 3 || id == 4 || id == 8 || id == 10;
  return combo && (bits & 15) == 0 && a.B_lo != nullptr;
}

int tc_gemm(const TcGemmArgs& a, cudaStream_t st) {
  if (!tc_gemm_supported(a)) { set_error("tc_gemm: unsupported shape/alignment (M=%lld N=%d K=%d)", a.M, a.N, a.K); return -2; }
  P p;
  p.A = a.A; p.lda = a.lda; p.M = a.M; p.N = a.N; p.K = a.K;
  plan(a.M, a.N, &p.BN, &p.n_tiles);
  p.m_tiles = (int)ceil_div(a.M, BM);
  p.k_blocks = (int)ceil_div(a.K, BK);
  const int stage_bytes = A_TILE + 2 * p.BN * 128;
  const int fixed = 1024 + 8 * STG_BYTES + 2 * 256 * 4 + 256;
  p.nstages = (SMEM_LIMIT - fixed) / stage_bytes;
  if (p.nstages > MAX_STAGES) p.nstages = MAX_STAGES;
  if (p.nstages < 2) { set_error("tc_gemm: not enough shared memory"); return -2; }
  const int smem_bytes = fixed + p.nstages * stage_bytes;
  p.bias = a.bias; p.relu = a.relu; p.gate = a.gate; p.gate_ld = a.gate_ld; p.gate_scale = a.gate_scale;
  p.drop_p = a.drop_p; p.rng = a.rng; p.drop_site = a.drop_site; p.resid = a.resid; p.resid_ld = a.resid_ld;
  p.drop_mask = a.drop_mask; p.drop_mask_ld = a.drop_mask_ld;
  p.dbg = g_gemm_dbg;

  CUtensorMap tmA, tmB, tmBlo, tmC;
  {
    cuuint64_t ad[2] = {(cuuint64_t)a.K, (cuuint64_t)a.M};
    cuuint64_t as_[1] = {(cuuint64_t)a.lda * 4};
    cuuint32_t ab[2] = {BK, BM};
    RD_TRY(encode(&tmA, a.A, 2, ad, as_, ab, CU_TENSOR_MAP_SWIZZLE_128B, "A"));
  }
  cuuint64_t bd[2] = {(cuuint64_t)a.K, (cuuint64_t)a.N};
  cuuint64_t bs[1] = {(cuuint64_t)a.K * 4};
  cuuint32_t bb[2] = {BK, (cuuint32_t)p.BN};
  RD_TRY(encode(&tmB, a.B, 2, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_128B, "B"));
  RD_TRY(encode(&tmBlo, a.B_lo, 2, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_128B, "B_lo"));
  cuuint64_t cd[2] = {(cuuint64_t)a.N, (cuuint64_t)a.M};
  cuuint64_t cs[1] = {(cuuint64_t)a.N * 4};
  cuuint32_t cb[2] = {32, 32};
  RD_TRY(encode(&tmC, a.C, 2, cd, cs, cb, CU_TENSOR_MAP_SWIZZLE_128B, "C"));
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  auto launch = [&](auto kern, int id) -> int {
    RD_TRY(ensure_attr((const void*)kern, id));
    launch_pdl(kern, dim3(grid), dim3(NTHREADS), smem_bytes, st, tmA, tmB, tmBlo, tmC, p);
    return 0;
  };
  const int id = (a.relu ? 8 : 0) | (a.gate ? 4 : 0) | (a.drop_p > 0.f ? 2 : 0) | (a.resid ? 1 : 0);
  int rc = -2;
  switch (id) {   // the combinations the encoder uses (forward: bias[,relu][,dropout][,residual]; backward: [gate][,residual])
    case 0: rc = launch(tc_gemm_kernel<false, false, false, false>, id); break;
    case 1: rc = launch(tc_gemm_kernel<false, false, false, true>, id); break;
    case 2: rc = launch(tc_gemm_kernel<false, false, true, false>, id); break;
    case 3: rc = launch(tc_gemm_kernel<false, false, true, true>, id); break;
    case 4: rc = launch(tc_gemm_kernel<false, true, false, false>, id); break;
    case 8: rc = launch(tc_gemm_kernel<true, false, false, false>, id); break;
    case 10: rc = launch(tc_gemm_kernel<true, false, true, false>, id); break;
    default: set_error("tc_gemm: epilogue combination %d not instantiated", id); return -2;
  }
  if (rc != 0) return rc;
  RD_CHECK_LAUNCH("tc_gemm_kernel");
  return 0;
}

bool tc_wgrad_supported(int Nout, int Kin, long long ldy, long long ldx, const void* dY, const void* X) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("RD_TC_WGRAD"); env = (e && e[0] == '0') ? 0 : 1; }
  if (env != 1) return false;
  if (Nout % 4 || Kin % 4 || ldy % 4 || ldx % 4 || Nout < 16 || Kin < 16) return false;
  return ((reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(X)) & 15) == 0;
}

long long tc_wgrad_partial_floats(int Nout, int Kin, long long rows) {
  WPlan w = wgrad_plan(Nout, Kin, rows);
  return round_up((long long)w.nsplit * w.Mpad * w.Nld, 64);
}

int tc_wgrad_group(const WgradItem* items, int n, const ColsumItem* cs, int ncs, cudaStream_t st) {
  if (n <= 0 && ncs <= 0) return 0;
  if (n > WG_MAX || ncs > CS_MAX) { set_error("tc_wgrad_group: at most %d problems (+ %d column sums) per launch", WG_MAX, CS_MAX); return -2; }
  WGroup g;
  RGroup r;
  g.n = n; r.n = n + (ncs > 0 ? ncs : 0);
  // rows per work item: the smallest target >= 512 for which the group's item count fills whole rounds of the grid
  auto count_items = [&](int target) {
    long long t = 0;
    for (int i = 0; i < n; ++i) { const WPlan w = wgrad_plan(items[i].Nout, items[i].Kin, items[i].rows, target); t += (long long)w.nsplit * w.m_tiles * w.n_tiles; }
    return t;
  };
  int target = 512;
  if (n > 0) {
    const long long sms = num_sms(), t0 = count_items(512);
    if (t0 > sms && t0 % sms) {
      const long long goal = (t0 / sms) * sms;
      for (int t = 544; t <= 1024; t += 32)
        if (count_items(t) <= goal) { target = t; break; }
    }
  }
  int item = 0, max_bn = 32;
  long long tot = 0;
  for (int i = 0; i < n; ++i) {
    const WgradItem& a = items[i];
    if (!tc_wgrad_supported(a.Nout, a.Kin, a.ldy, a.ldx, a.dY, a.X) || !a.partial || a.rows < 1 ||
        (reinterpret_cast<uintptr_t>(a.partial) & 15)) {
      set_error("tc_wgrad_group: problem %d has an unsupported shape/alignment", i);
      return -2;
    }
    const WPlan w = wgrad_plan(a.Nout, a.Kin, a.rows, target);
    WP& p = g.it[i];
    p.partial = a.partial;
    p.rows = a.rows; p.M = a.Nout; p.N = a.Kin;
    {
      cuuint32_t box[2] = {32, BK};
      cuuint64_t da[2] = {(cuuint64_t)a.Nout, (cuuint64_t)a.rows}, sa_[1] = {(cuuint64_t)a.ldy * 4};
      RD_TRY(encode(&g.tmA[i], a.dY, 2, da, sa_, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "wgrad dY"));
      cuuint64_t db_[2] = {(cuuint64_t)a.Kin, (cuuint64_t)a.rows}, sb_[1] = {(cuuint64_t)a.ldx * 4};
      RD_TRY(encode(&g.tmB[i], a.X, 2, db_, sb_, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "wgrad X"));
    }
    p.BN = w.BN; p.n_tiles = w.n_tiles; p.m_tiles = w.m_tiles; p.nsplit = w.nsplit; p.rows_per_split = w.rows_per_split;
    p.Mpad = w.Mpad; p.Nld = w.Nld;
    if (p.BN > max_bn) max_bn = p.BN;
    if (a.rows > 0x7fffffffLL) { set_error("tc_wgrad_group: too many rows"); return -2; }
    p.item0 = item;
    item += w.nsplit * w.m_tiles * w.n_tiles;
    RItem& q = r.it[i];
    q.partial = a.partial; q.dW = a.dW; q.db = a.db; q.nsplit = w.nsplit; q.M = a.Nout; q.N = a.Kin; q.Mpad = w.Mpad; q.Nld = w.Nld;
    q.kind = 0; q.stride = 0;
    q.start = tot;
    tot += (long long)a.Nout * (a.Kin / 4 + 1);
  }
  g.total_items = item;
  g.stage_bytes = 2 * A_TILE + 2 * max_bn * 128;
  const int fixed = 1024 + 256;
  g.nstages = (SMEM_LIMIT - fixed) / g.stage_bytes;
  if (g.nstages > MAX_STAGES) g.nstages = MAX_STAGES;
  if (n > 0 && g.nstages < 2) { set_error("tc_wgrad_group: not enough shared memory"); return -2; }
  const int smem_bytes = fixed + g.nstages * g.stage_bytes;
  for (int i = 0; i < ncs; ++i) {
    RItem& q = r.it[n + i];
    q.partial = cs[i].partial; q.dW = cs[i].out; q.db = nullptr; q.nsplit = cs[i].nsplit; q.M = 1; q.N = cs[i].ncols;
    q.Mpad = 1; q.Nld = 0; q.kind = 1; q.stride = cs[i].stride;
    tot = round_up(tot, 32);                 // warp-aligned: 32 threads per output column
    q.start = tot;
    tot += 32LL * cs[i].ncols;
  }
  r.total = tot;
  if (n > 0) {
    RD_TRY(ensure_attr((const void*)tc_wgrad_kernel, 15));
    const int grid = item < num_sms() ? item : num_sms();
    launch_pdl(tc_wgrad_kernel, dim3(grid), dim3(W_THREADS), smem_bytes, st, g);
    RD_CHECK_LAUNCH("tc_wgrad_kernel");
  }
  launch_pdl(wgrad_reduce_kernel, dim3((unsigned)ceil_div(tot, 256)), dim3(256), 0, st, r);
  RD_CHECK_LAUNCH("wgrad_reduce_kernel");
  return 0;
}

int tc_wgrad(const float* dY, long long ldy, const float* X, long long ldx, long long rows, int Nout, int Kin,
             float* dW, float* db, float* partial, cudaStream_t st) {
  WgradItem it{dY, ldy, X, ldx, rows, Nout, Kin, dW, db, partial};
  return tc_wgrad_group(&it, 1, nullptr, 0, st);
}

int split_weights(const WeightSplit* items, int n, cudaStream_t st, const StepPrologue* pro) {
  if (n <= 0 && !pro) return 0;
  if (n > 16) { set_error("split_weights: at most 16 tensors per launch"); return -2; }
  SplitItems s;
  s.n = n; s.start[0] = 0;
  for (int i = 0; i < n; ++i) { s.it[i] = items[i]; s.start[i + 1] = s.start[i] + (long long)items[i].rows * items[i].cols; }
  s.pro = pro ? *pro : StepPrologue{};
  const long long total = s.start[n] > 0 ? s.start[n] : 1;
  launch_pdl(split_weights_kernel, dim3((unsigned)ceil_div(total, 256)), dim3(256), 0, st, s);
  RD_CHECK_LAUNCH("split_weights_kernel");
  return 0;
}

}  // namespace rd
