// yy_gemm.cuh -- interface of the tcgen05 GEMM building block (yy_gemm.cu).
#pragma once
#include <cuda_bf16.h>
#include "yy_common.cuh"

namespace yy {

struct GemmArgs {
  const __nv_bfloat16* A; int lda;  // [M][K] row-major
  const __nv_bfloat16* B; int ldb;  // [N][K] row-major (nn.Linear weight layout)
  float* C; int ldc;                // [M][N] fp32
  const float* bias;                // [N] or nullptr
  int M, N, K;
  int relu;
};

// C = act(A * B^T + bias).  N multiple of 16; K, lda, ldb multiples of 8.
int gemm_bf16_tn(const GemmArgs& g, cudaStream_t stream);
// Two independent problems with the same M in ONE launch (policy FC + value FC1 of the heads): grid.y covers
// p0's N tiles then p1's.  Both N multiples of 64.
int gemm_bf16_tn_pair(const GemmArgs& p0, const GemmArgs& p1, cudaStream_t stream);

}  // namespace yy
