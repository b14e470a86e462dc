// Classification head of Raindrop_v2 as three small fused kernels (code/models_rd.py:293-294,383-385):
//   feat = [pooled || emb(static)],  h = relu(mlp_static.0(feat)),  logits = mlp_static.2(h)
// At the reference's batch sizes these are 128 x 186 problems: eleven generic GEMM/reduce launches of
// ~10 us each were pure latency.  One CTA per sample for the forward and the per-sample backward; the
// weight gradients are thread-per-element sums over the batch in a fixed order (deterministic).
#include "rd_kernels.cuh"

namespace rd {
namespace {

constexpr int HT = 256;

struct HeadP {
  int B, D, N, ds, Df, ncls;
  const float* statics; const float* emb_w; const float* emb_b;
  const float* w0; const float* b0; const float* w2; const float* b2;
};

// grid = B.  feat[b, :D] already holds the pooled encoder output (masked_mean_fwd).
__global__ void __launch_bounds__(HT) head_fwd_kernel(HeadP p, float* __restrict__ feat, float* __restrict__ hpre,
                                                      float* __restrict__ logits) {
  extern __shared__ float sm[];
  float* fs = sm; float* hs = sm + p.Df;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* fb = feat + (long long)b * p.Df;
  for (int k = tid; k < p.D; k += HT) fs[k] = fb[k];
  if (p.ds > 0) {   // emb = Linear(d_static, N)(static)                      code/models_rd.py:293-294
    for (int n = tid; n < p.N; n += HT) {
      float a = __ldg(p.emb_b + n);
      for (int k = 0; k < p.ds; ++k) a = fmaf(__ldg(p.statics + (long long)b * p.ds + k), __ldg(p.emb_w + n * p.ds + k), a);
      fs[p.D + n] = a;
      fb[p.D + n] = a;
    }
  }
  __syncthreads();
  // a warp owns 4 hidden units at a time (4 independent coalesced weight-row streams in flight), shuffle reduce
  for (int j0 = warp * 4; j0 < p.Df; j0 += (HT / 32) * 4) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < p.Df; k += 32) {
      const float f = fs[k];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + u < p.Df) a[u] = fmaf(f, __ldg(p.w0 + (long long)(j0 + u) * p.Df + k), a[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float v = warp_sum(a[u]);
      if (lane == 0 && j0 + u < p.Df) {
        const float h = fmaxf(v + __ldg(p.b0 + j0 + u), 0.f);
        hs[j0 + u] = h;
        hpre[(long long)b * p.Df + j0 + u] = h;
      }
    }
  }
  __syncthreads();
  for (int c = warp; c < p.ncls; c += HT / 32) {
    const float* wr = p.w2 + (long long)c * p.Df;
    float a = 0.f;
    for (int j = lane; j < p.Df; j += 32) a = fmaf(hs[j], __ldg(wr + j), a);
    a = warp_sum(a);
    if (lane == 0) logits[(long long)b * p.ncls + c] = a + __ldg(p.b2 + c);
  }
}

// grid = B: dh = (dlogits . W2) * [h > 0];  dfeat = dh . W0
__global__ void __launch_bounds__(HT) head_bwd_sample_kernel(HeadP p, const float* __restrict__ hpre,
                                                             const float* __restrict__ dlogits, float* __restrict__ dh,
                                                             float* __restrict__ dfeat) {
  extern __shared__ float sm[];
  float* ds_ = sm;   // dh of this sample
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int j = tid; j < p.Df; j += HT) {
    float a = 0.f;
    for (int c = 0; c < p.ncls; ++c) a = fmaf(__ldg(dlogits + (long long)b * p.ncls + c), __ldg(p.w2 + (long long)c * p.Df + j), a);
    a = hpre[(long long)b * p.Df + j] > 0.f ? a : 0.f;
    ds_[j] = a;
    dh[(long long)b * p.Df + j] = a;
  }
  __syncthreads();
  for (int k = tid; k < p.Df; k += HT) {          // column k of W0: coalesced across threads
    float a = 0.f;
#pragma unroll 8
    for (int j = 0; j < p.Df; ++j) a = fmaf(ds_[j], __ldg(p.w0 + (long long)j * p.Df + k), a);
    dfeat[(long long)b * p.Df + k] = a;
  }
}

// out[j, k] = sum_b L[b, j] * R[b, k]  (j < J, k < K), bias[j] = sum_b L[b, j]; block (32 k, 8 j)
__global__ void head_outer_kernel(const float* __restrict__ Lm, long long ldl, const float* __restrict__ Rm, long long ldr,
                                  int B, int J, int K, float* __restrict__ out, float* __restrict__ bias) {
  const int k = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
  if (j >= J) return;
  float a = 0.f, s = 0.f;
  const bool kin = k < K;
  const int kk = kin ? k : 0;                 // out-of-range lanes read a valid column and discard the result
#pragma unroll 8
  for (int b = 0; b < B; ++b) {                // 8 independent load pairs in flight per thread, fixed summation order
    const float l = __ldg(Lm + (long long)b * ldl + j);
    s += l;
    a = fmaf(l, __ldg(Rm + (long long)b * ldr + kk), a);
  }
  if (kin) out[(long long)j * K + k] = a;
  if (bias && k == 0) bias[j] = s;
}

HeadP make(int B, int D, int N, int ds, int ncls, const float* statics, const float* emb_w, const float* emb_b,
           const float* w0, const float* b0, const float* w2, const float* b2) {
  HeadP p;
  p.B = B; p.D = D; p.N = N; p.ds = ds; p.Df = D + (ds > 0 ? N : 0); p.ncls = ncls;
  p.statics = statics; p.emb_w = emb_w; p.emb_b = emb_b; p.w0 = w0; p.b0 = b0; p.w2 = w2; p.b2 = b2;
  return p;
}

}  // namespace

int head_fwd(int B, int D, int N, int ds, int ncls, const float* statics, const float* emb_w, const float* emb_b,
             const float* w0, const float* b0, const float* w2, const float* b2, float* feat, float* hpre, float* logits,
             cudaStream_t st) {
  HeadP p = make(B, D, N, ds, ncls, statics, emb_w, emb_b, w0, b0, w2, b2);
  if (2 * p.Df * sizeof(float) > 48 * 1024) { set_error("head_fwd: feature width %d too large", p.Df); return -2; }
  head_fwd_kernel<<<B, HT, 2 * p.Df * sizeof(float), st>>>(p, feat, hpre, logits);
  RD_CHECK_LAUNCH("head_fwd_kernel");
  return 0;
}

int head_bwd(int B, int D, int N, int ds, int ncls, const float* statics, const float* w0, const float* w2,
             const float* feat, const float* hpre, const float* dlogits, float* dh, float* dfeat, float* g_w0, float* g_b0,
             float* g_w2, float* g_b2, float* g_emb_w, float* g_emb_b, cudaStream_t st) {
  HeadP p = make(B, D, N, ds, ncls, statics, nullptr, nullptr, w0, nullptr, w2, nullptr);
  head_bwd_sample_kernel<<<B, HT, p.Df * sizeof(float), st>>>(p, hpre, dlogits, dh, dfeat);
  RD_CHECK_LAUNCH("head_bwd_sample_kernel");
  const dim3 blk(32, 8);
  head_outer_kernel<<<dim3((unsigned)ceil_div(p.Df, 32), (unsigned)ceil_div(p.Df, 8)), blk, 0, st>>>(
      dh, p.Df, feat, p.Df, B, p.Df, p.Df, g_w0, g_b0);                       // d mlp_static.0
  RD_CHECK_LAUNCH("head_outer_kernel");
  head_outer_kernel<<<dim3((unsigned)ceil_div(p.Df, 32), (unsigned)ceil_div(ncls, 8)), blk, 0, st>>>(
      dlogits, ncls, hpre, p.Df, B, ncls, p.Df, g_w2, g_b2);                  // d mlp_static.2
  RD_CHECK_LAUNCH("head_outer_kernel");
  if (ds > 0) {
    head_outer_kernel<<<dim3((unsigned)ceil_div(ds, 32), (unsigned)ceil_div(N, 8)), blk, 0, st>>>(
        dfeat + D, p.Df, statics, ds, B, N, ds, g_emb_w, g_emb_b);            // d emb
    RD_CHECK_LAUNCH("head_outer_kernel");
  }
  return 0;
}

}  // namespace rd
