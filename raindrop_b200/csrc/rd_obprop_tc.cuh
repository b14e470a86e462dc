// tcgen05 / TMEM / TMA kernel for one observation-propagation layer (rd_obprop_tc.cu).
#pragma once
#include "rd_common.cuh"

namespace rd {

// true when the tensor-core kernel handles this layer shape (C = T*d_ob channels)
bool obprop_tc_supported(int C);

// out[r, :] = relu(x[r, :] . W^T + b) * scale[r % mod], TF32 operands, fp32 accumulate in TMEM.
// perm != 0: store into the encoder input [T, B, D] instead of [rows, C] (needs d_ob == 4):
//   row r = b*pN + n, col c = t*4 + k  ->  out[((t*pB + b)*pD) + n*4 + k]
int obprop_tc_fwd(const float* x, const float* W, const float* b, const float* scale, int mod, int64_t rows,
                  int C, float* out, int perm, int pB, int pN, int pdob, int pD, cudaStream_t st);

}  // namespace rd
