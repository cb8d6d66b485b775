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

}  // namespace yy
