// Fused temporal self-attention for short sequences (T <= 64, head_dim <= 96): one CTA per (sample, head)
// keeps Q, K, V [T x hd] and the T x T score tile in shared memory, so QK^T, the key-padding-masked
// softmax, attention dropout and PV are one launch (forward) and the whole backward is one launch
// (probabilities are RECOMPUTED from Q, K and the counter-based dropout mask: nothing T x T is stored).
// This is the temporal-attention stage of nn.TransformerEncoder as called at code/models_rd.py:358 for
// the P19 shape (T = 60, hd = 76).  A 60 x 60 x 76 problem is far below one 128-row UMMA tile and needs
// fp32 accuracy, so it runs on the CUDA cores with 4x4 / 4x5 register tiles; longer sequences take the
// batched-GEMM path (rd_model.cu).
#include "rd_kernels.cuh"

namespace rd {
namespace {

constexpr int TM = 64;      // max sequence length
constexpr int HDM = 96;     // max head dim
constexpr int NT = 256;

struct AttnP {
  const float* qkv; float* ctx;            // forward
  const float* dctx; float* dqkv;          // backward
  const int64_t* lengths;
  int B, H, T, hd, D;
  float scale, drop_p;
  const uint64_t* rng; uint32_t site;
};

// S = scale * Q K^T with key-padding mask, softmax -> Ps (probabilities) and Pd (dropped copy)
__device__ __forceinline__ void scores_softmax(const AttnP& p, int b, int h, const float* Qs, const float* Ks, float* Ps,
                                               float* Pd, int ldq) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int T = p.T, hd = p.hd;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  for (int d = 0; d < hd; ++d) {
    float q[4], k[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) q[r] = Qs[(4 * ty + r) * ldq + d];
#pragma unroll
    for (int c = 0; c < 4; ++c) k[c] = Ks[(4 * tx + c) * ldq + d];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(q[r], k[c], acc[r][c]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) Ps[(4 * ty + r) * (TM + 1) + 4 * tx + c] = acc[r][c] * p.scale;
  __syncthreads();
  const long long len = p.lengths[b];
  const int nv = (int)(len < T ? (len < 0 ? 0 : len) : T);
  const float ik = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < T; i += NT / 32) {
    float* row = Ps + i * (TM + 1);
    float v0 = lane < nv ? row[lane] : -INFINITY, v1 = lane + 32 < nv ? row[lane + 32] : -INFINITY;
    float mx = warp_max(fmaxf(v0, v1));
    float e0 = lane < nv ? expf(v0 - mx) : 0.f, e1 = lane + 32 < nv ? expf(v1 - mx) : 0.f;
    float sum = warp_sum(e0 + e1);
    float inv = nv > 0 ? 1.f / sum : 0.f;
    const uint64_t base = ((uint64_t)(b * p.H + h) * T + i) * T;     // index space [B, H, T, T]
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      int j = lane + 32 * half;
      if (j < TM) {
        float pr = (half ? e1 : e0) * inv;
        row[j] = pr;
        float m = (p.drop_p > 0.f && j < T) ? dropout_scale(p.rng, p.site, base + j, p.drop_p, ik) : 1.f;
        Pd[i * (TM + 1) + j] = pr * m;
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void load_head(const float* __restrict__ src, long long row_stride, int T, int hd, float* dst,
                                          int ld) {
  for (int idx = threadIdx.x; idx < TM * ld; idx += NT) {
    int t = idx / ld, d = idx - t * ld;
    dst[idx] = (t < T && d < hd) ? __ldg(src + (long long)t * row_stride + d) : 0.f;
  }
}

__global__ void __launch_bounds__(NT) attn_small_fwd_kernel(AttnP p) {
  extern __shared__ float sm[];
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int ld = p.hd + 1;
  float* Qs = sm; float* Ks = Qs + TM * ld; float* Vs = Ks + TM * ld;
  float* Ps = Vs + TM * ld; float* Pd = Ps + TM * (TM + 1);
  const long long rs = (long long)p.B * 3 * p.D;
  const float* base = p.qkv + (long long)b * 3 * p.D + h * p.hd;
  load_head(base, rs, p.T, p.hd, Qs, ld);
  load_head(base + p.D, rs, p.T, p.hd, Ks, ld);
  load_head(base + 2 * p.D, rs, p.T, p.hd, Vs, ld);
  __syncthreads();
  scores_softmax(p, b, h, Qs, Ks, Ps, Pd, ld);
  // ctx[i, d] = sum_j Pd[i, j] V[j, d]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][6];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[r][c] = 0.f;
  for (int j = 0; j < p.T; ++j) {
    float pv[4], vv[6];
#pragma unroll
    for (int r = 0; r < 4; ++r) pv[r] = Pd[(4 * ty + r) * (TM + 1) + j];
#pragma unroll
    for (int c = 0; c < 6; ++c) vv[c] = Vs[j * ld + tx + 16 * c];     // columns >= hd are zero padding or unused
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[r][c] = fmaf(pv[r], vv[c], acc[r][c]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int i = 4 * ty + r;
    if (i >= p.T) continue;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      int d = tx + 16 * c;
      if (d < p.hd) p.ctx[((long long)i * p.B + b) * p.D + h * p.hd + d] = acc[r][c];
    }
  }
}

__global__ void __launch_bounds__(NT) attn_small_bwd_kernel(AttnP p) {
  extern __shared__ float sm[];
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int ld = p.hd + 1;
  float* Qs = sm; float* Ks = Qs + TM * ld; float* Vs = Ks + TM * ld; float* Gs = Vs + TM * ld;   // Gs = d(ctx)
  float* Ps = Gs + TM * ld; float* Pd = Ps + TM * (TM + 1); float* dS = Pd + TM * (TM + 1);
  const long long rs = (long long)p.B * 3 * p.D;
  const float* base = p.qkv + (long long)b * 3 * p.D + h * p.hd;
  load_head(base, rs, p.T, p.hd, Qs, ld);
  load_head(base + p.D, rs, p.T, p.hd, Ks, ld);
  load_head(base + 2 * p.D, rs, p.T, p.hd, Vs, ld);
  load_head(p.dctx + (long long)b * p.D + h * p.hd, (long long)p.B * p.D, p.T, p.hd, Gs, ld);
  __syncthreads();
  scores_softmax(p, b, h, Qs, Ks, Ps, Pd, ld);            // recompute P and the dropped copy
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  {  // dPd[i, j] = sum_d G[i, d] V[j, d]
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    for (int d = 0; d < p.hd; ++d) {
      float g[4], v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) g[r] = Gs[(4 * ty + r) * ld + d];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = Vs[(4 * tx + c) * ld + d];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(g[r], v[c], acc[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) dS[(4 * ty + r) * (TM + 1) + 4 * tx + c] = acc[r][c];
  }
  __syncthreads();
  {  // dS = P * (dP - rowsum(dP * P)),  dP = dPd * mask/(1-p) = dPd * Pd / P where P > 0
    const int warp = tid >> 5, lane = tid & 31;
    const float ik = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    for (int i = warp; i < p.T; i += NT / 32) {
      const uint64_t ibase = ((uint64_t)(b * p.H + h) * p.T + i) * p.T;
      float dp[2], pr[2];
      float dot = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        int j = lane + 32 * half;
        pr[half] = j < p.T ? Ps[i * (TM + 1) + j] : 0.f;
        float m = (p.drop_p > 0.f && j < p.T) ? dropout_scale(p.rng, p.site, ibase + j, p.drop_p, ik) : 1.f;
        dp[half] = j < p.T ? dS[i * (TM + 1) + j] * m : 0.f;
        dot += dp[half] * pr[half];
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        int j = lane + 32 * half;
        if (j < TM) dS[i * (TM + 1) + j] = pr[half] * (dp[half] - dot);
      }
    }
  }
  __syncthreads();
  // dQ[i, d] = scale * sum_j dS[i, j] K[j, d];  dK[j, d] = scale * sum_i dS[i, j] Q[i, d];  dV[j, d] = sum_i Pd[i, j] G[i, d]
  float aq[4][6], ak[4][6], av[4][6];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) { aq[r][c] = 0.f; ak[r][c] = 0.f; av[r][c] = 0.f; }
  for (int j = 0; j < p.T; ++j) {
    float s_row[4], s_col[4], p_col[4], kk[6], qq[6], gg[6];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      s_row[r] = dS[(4 * ty + r) * (TM + 1) + j];      // dS[i = 4ty+r, j]
      s_col[r] = dS[j * (TM + 1) + 4 * ty + r];        // dS[i = j, j' = 4ty+r]  (transposed use)
      p_col[r] = Pd[j * (TM + 1) + 4 * ty + r];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) { int d = tx + 16 * c; kk[c] = Ks[j * ld + d]; qq[c] = Qs[j * ld + d]; gg[c] = Gs[j * ld + d]; }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        aq[r][c] = fmaf(s_row[r], kk[c], aq[r][c]);
        ak[r][c] = fmaf(s_col[r], qq[c], ak[r][c]);
        av[r][c] = fmaf(p_col[r], gg[c], av[r][c]);
      }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int i = 4 * ty + r;
    if (i >= p.T) continue;
    float* o = p.dqkv + ((long long)i * p.B + b) * 3 * p.D + h * p.hd;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      int d = tx + 16 * c;
      if (d < p.hd) { o[d] = aq[r][c] * p.scale; o[p.D + d] = ak[r][c] * p.scale; o[2 * p.D + d] = av[r][c]; }
    }
  }
}

size_t fwd_smem(int hd) { return sizeof(float) * (3 * TM * (hd + 1) + 2 * TM * (TM + 1)); }
size_t bwd_smem(int hd) { return sizeof(float) * (4 * TM * (hd + 1) + 3 * TM * (TM + 1)); }

}  // namespace

bool attn_small_supported(int T, int hd) {
  static int env = -1;
  if (env < 0) { const char* e = getenv("RD_ATTN_SMALL"); env = (e && e[0] == '0') ? 0 : 1; }
  return env == 1 && T <= TM && hd <= HDM;
}

int attn_small_fwd(const float* qkv, const int64_t* lengths, int B, int H, int T, int hd, float drop_p,
                   const uint64_t* rng, uint32_t site, float* ctx, cudaStream_t st) {
  AttnP p{};
  p.qkv = qkv; p.ctx = ctx; p.lengths = lengths; p.B = B; p.H = H; p.T = T; p.hd = hd; p.D = H * hd;
  p.scale = 1.f / sqrtf((float)hd); p.drop_p = drop_p; p.rng = rng; p.site = site;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(attn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem(HDM));
    cudaFuncSetAttribute(attn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem(HDM));
    attr = true;
  }
  attn_small_fwd_kernel<<<B * H, NT, fwd_smem(hd), st>>>(p);
  RD_CHECK_LAUNCH("attn_small_fwd_kernel");
  return 0;
}

int attn_small_bwd(const float* qkv, const float* dctx, const int64_t* lengths, int B, int H, int T, int hd, float drop_p,
                   const uint64_t* rng, uint32_t site, float* dqkv, cudaStream_t st) {
  AttnP p{};
  p.qkv = qkv; p.dctx = dctx; p.dqkv = dqkv; p.lengths = lengths; p.B = B; p.H = H; p.T = T; p.hd = hd; p.D = H * hd;
  p.scale = 1.f / sqrtf((float)hd); p.drop_p = drop_p; p.rng = rng; p.site = site;
  attn_small_bwd_kernel<<<B * H, NT, bwd_smem(hd), st>>>(p);
  RD_CHECK_LAUNCH("attn_small_bwd_kernel");
  return 0;
}

}  // namespace rd
