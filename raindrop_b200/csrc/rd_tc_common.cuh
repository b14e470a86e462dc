// Shared device-side PTX wrappers (mbarrier, TMA, tcgen05, TMEM) and host-side tensor-map helpers of
// the tensor-core kernels (rd_obprop_tc.cu, rd_tc_gemm.cu).  sm_100a only.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include <mutex>

#include "rd_common.cuh"

namespace rd {
namespace tc {

constexpr int SMEM_LIMIT = 232448;   // 227 KB of dynamic shared memory per CTA

__device__ __forceinline__ float rn_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// A operand taken from tensor memory (lane = row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K-major operand tile, rows of 128 bytes, 128B swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}


// MN-major TF32 operand tile (the contraction index runs over the ROWS of a row-major tile).  For 32-bit operands the
// tensor core accepts exactly one MN-major shared-memory layout: 128-byte lines of 32 consecutive M/N elements whose
// 32-BYTE chunks are XOR-swizzled with (line & 3) -- Swizzle<2,5,2>, descriptor layout type SWIZZLE_128B_BASE32B,
// which is what TMA produces with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  K atoms are 4 lines (SBO = 512 bytes),
// 32-element M/N groups are `lbo_bytes` apart (LBO).  One kind::tf32 MMA (K = 8) consumes 8 lines = 1024 bytes.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}
// byte offset of the 16-byte piece holding columns 4*q4 .. 4*q4+3 (q4 < 8) of line `r` inside one 32-column group
__device__ __forceinline__ uint32_t mn_sw_offset(int r, int q4) {
  return (uint32_t)(r * 128 + ((((q4 >> 1) ^ (r & 3))) << 5) + ((q4 & 1) << 4));
}
// instruction descriptor, kind::tf32, fp32 accumulate; a_mn / b_mn: operand is MN-major
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Remainder images for the error-compensated ("3xTF32") products: dst[i] = x - trunc19(x) for `nvec` 16-byte vectors
// starting at shared address `src`, vector index = tid + k * nthreads.  All loads of a batch of NB vectors are issued
// before the first store, so a thread has NB shared-memory round trips in flight instead of one.
template <int NB>
__device__ __forceinline__ void lo_image(uint32_t src, uint32_t dst, uint32_t nvec, uint32_t tid, uint32_t nthreads) {
  for (uint32_t v0 = tid; v0 < nvec; v0 += NB * nthreads) {
    float4 x[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const uint32_t v = v0 + k * nthreads;
      if (v < nvec)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[k].x), "=f"(x[k].y), "=f"(x[k].z), "=f"(x[k].w) : "r"(src + v * 16u));
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const uint32_t v = v0 + k * nthreads;
      if (v < nvec) {
        const float a = x[k].x - __uint_as_float(__float_as_uint(x[k].x) & 0xFFFFE000u);
        const float b = x[k].y - __uint_as_float(__float_as_uint(x[k].y) & 0xFFFFE000u);
        const float c = x[k].z - __uint_as_float(__float_as_uint(x[k].z) & 0xFFFFE000u);
        const float d = x[k].w - __uint_as_float(__float_as_uint(x[k].w) & 0xFFFFE000u);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + v * 16u), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
      }
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// cuTensorMapEncodeTiled costs a few microseconds of host time; a training step re-encodes the same
// ~90 maps every step (same buffers, same shapes), so keep them in a small direct-mapped cache.
struct TmapKey {
  const void* addr; int rank; int swizzle; cuuint64_t dims[5]; cuuint64_t strides[4]; cuuint32_t box[5];
  bool operator==(const TmapKey& o) const {
    if (addr != o.addr || rank != o.rank || swizzle != o.swizzle) return false;
    for (int i = 0; i < 5; ++i) if (dims[i] != o.dims[i] || box[i] != o.box[i]) return false;
    for (int i = 0; i < 4; ++i) if (strides[i] != o.strides[i]) return false;
    return true;
  }
};
struct TmapSlot { bool valid = false; TmapKey key; CUtensorMap map; };

inline int encode(CUtensorMap* m, const void* addr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, CUtensorMapSwizzle sw, const char* what) {
  constexpr int NSLOT = 512;
  static thread_local TmapSlot cache[NSLOT];
  TmapKey k{};
  k.addr = addr; k.rank = rank; k.swizzle = (int)sw;
  for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) k.strides[i] = strides_bytes[i];
  uint64_t h = reinterpret_cast<uintptr_t>(addr) * 0x9E3779B97F4A7C15ull;
  h ^= (dims[0] * 0xC2B2AE3D27D4EB4Full) ^ ((uint64_t)box[rank - 1] << 17) ^ (rank > 1 ? dims[1] * 0x165667B19E3779F9ull : 0);
  TmapSlot& slot = cache[(h >> 40) % NSLOT];
  if (slot.valid && slot.key == k) { *m = slot.map; return 0; }
  EncodeTiledFn fn = get_encode();
  if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return -3; }
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(addr), dims, strides_bytes,
                  box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed: CUresult %d", what, (int)r); return -3; }
  slot.valid = true; slot.key = k; slot.map = *m;
  return 0;
}

// cudaFuncSetAttribute(max dynamic shared memory) once per (kernel, device): a process that touches a second
// GPU has to opt in there as well (the attribute is per device).
inline int ensure_max_smem(const void* fn, int bytes) {
  struct Slot { const void* fn; int dev; };
  static Slot done[256];
  static int n = 0;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n; ++i) if (done[i].fn == fn && done[i].dev == dev) return 0;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem): %s", cudaGetErrorString(e)); return -1; }
  if (n < 256) { done[n].fn = fn; done[n].dev = dev; ++n; }
  return 0;
}

inline int num_sms() {
  static int n[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = (dev >= 0 && dev < 16) ? dev : 0;
  if (n[slot] == 0) {
    cudaDeviceGetAttribute(&n[slot], cudaDevAttrMultiProcessorCount, dev);
    if (n[slot] <= 0) n[slot] = 148;
  }
  return n[slot];
}

}  // namespace tc
}  // namespace rd
