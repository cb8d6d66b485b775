// yy_probe.cu -- tcgen05 self-tests and micro-benchmarks (developer tools; nothing here is on the product path).
//   yy_probe_umma : one-CTA C[128,N] = A[M,K] * B[N,K]^T with the no-swizzle K-major core-matrix layout [k/8][row][8]
//                   and descriptor strides the persistent kernel uses (LBO = rows*16, SBO = 128 or a board-row pitch);
//                   tests/ pin the descriptor conventions with it, including start addresses a 3x3 tap shift produces.
//   yy_umma_rate  : cycles per tcgen05.mma vs shared-memory layout (operand-fetch rate: 128 B/clk shared-memory port).
//   yy_l2_stream  : achievable L2 -> shared-memory bulk-copy rate when every SM streams the same weight image.
#include <cuda_bf16.h>

#include "yy_common.cuh"
#include "yy_ptx.cuh"

namespace yy {
using namespace ptx;

// ------------------------------------------------------------------------------------------------ probe
// Single CTA, everything resident: C[128,N] = A[row_off : row_off+128, :] * B^T.  rowsA >= row_off + 128.
__global__ void __launch_bounds__(128) probe_umma_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                        float* __restrict__ C, int rowsA, int N, int K, int row_off, int swap) {
  // swap >= 2: 'strided-group' mode -- the MMA's 8-row groups start every `swap` rows of A (SBO = swap*16 B), i.e.
  // C row m = A[row_off + (m/8)*swap + m%8]: what the tower kernel's row-aligned layout relies on (SBO = 9*16).
  const int grp = swap >= 2 ? swap : 8;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int KC = K / 8;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + (size_t)KC * rowsA * 16;
  for (int i = threadIdx.x; i < KC * rowsA; i += blockDim.x) {
    int kc = i / rowsA, r = i % rowsA;
    *reinterpret_cast<uint4*>(a_s + (size_t)i * 16) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + kc * 8);
  }
  for (int i = threadIdx.x; i < KC * N; i += blockDim.x) {
    int kc = i / N, r = i % N;
    *reinterpret_cast<uint4*>(b_s + (size_t)i * 16) = *reinterpret_cast<const uint4*>(B + (size_t)r * K + kc * 8);
  }
  uint32_t ncols = 32; while ((int)ncols < N) ncols <<= 1;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), ncols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_bf16(128, N);
    for (int k16 = 0; k16 < K / 16; ++k16) {
      uint32_t a_addr = smem_u32(a_s) + (uint32_t)((2 * k16 * rowsA + row_off) * 16);
      uint32_t b_addr = smem_u32(b_s) + (uint32_t)(2 * k16 * N * 16);
      uint64_t ad = swap == 1 ? smem_desc(a_addr, 128, rowsA * 16) : smem_desc(a_addr, rowsA * 16, grp * 16);
      uint64_t bd = swap == 1 ? smem_desc(b_addr, 128, N * 16) : smem_desc(b_addr, N * 16, 128);
      tc_mma_bf16(tmem_base, ad, bd, idesc, k16 > 0 ? 1u : 0u);
    }
    tc_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = warp * 32 + lane;
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
    tc_wait_ld();
    for (int j = 0; j < 16; ++j) C[(size_t)row * N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, ncols);
}

}  // namespace yy

extern "C" int yy_probe_umma(const void* a, const void* b, float* c, int M, int N, int K, int a_row_offset,
                             int swap_lbo_sbo, void* stream) {
  using namespace yy;
  if (yy_device_count() == 0) return set_error(YY_ERR_NO_DEVICE, "no CUDA device");
  if (M < (swap_lbo_sbo >= 2 ? 15 * swap_lbo_sbo + 8 : 128) + a_row_offset || a_row_offset < 0) return set_error(YY_ERR_INVALID, "probe: A has too few rows");
  if (N < 16 || N > 256 || N % 16 || K < 16 || K % 16) return set_error(YY_ERR_INVALID, "probe: N in [16,256] step 16, K multiple of 16");
  size_t smem = (size_t)(K / 8) * 16 * ((size_t)M + N);
  if (smem > 200 * 1024) return set_error(YY_ERR_INVALID, "probe: operands too large for one CTA");
  YY_CUDA_OK(cudaFuncSetAttribute(probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, c, M, N, K,
                                                            a_row_offset, swap_lbo_sbo);
  YY_LAUNCH_CHECK();
  return YY_OK;
}

// ------------------------------------------------------------------------------------------------ UMMA rate bench
// Developer tool: issues `iters` rounds of `per_round` back-to-back tcgen05.mma (M=128, N, K=16) from one thread on
// zero-filled shared memory, with caller-chosen descriptor strides / swizzle mode, on every SM, and reports the
// SM cycles per MMA.  Used to choose the shared-memory layouts of the tower kernel (operand-fetch rate).
namespace yy {
__global__ void __launch_bounds__(128) umma_rate_kernel(int N, int layout_type, uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b,
                                                       uint32_t sbo_b, uint32_t a_step, uint32_t b_step, int per_round, int iters,
                                                       long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x < 32) {   // whole warp runs the loop; one elected lane issues (the pattern the tower kernel uses)
    const uint32_t idesc = idesc_bf16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 150 * 1024;
    const uint64_t lt = (uint64_t)layout_type << 61;
    const uint64_t ad0 = smem_desc(a0, lbo_a, sbo_a) | lt, bd0 = smem_desc(b0, lbo_b, sbo_b) | lt;
    const uint64_t da = a_step >> 4, db = b_step >> 4;
    const uint32_t dcol = N > 128 ? 0u : 128u;
    uint32_t phase = 0;
    long long total = 0;
    for (int it = 0; it < iters; ++it) {
      long long t0 = clock64();
      for (int i = 0; i < per_round; i += 16) {
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < 16; ++u)
            tc_mma_bf16(tmem_base + (uint32_t)(u & 3) * dcol, ad0 + (uint64_t)(u & 7) * da, bd0 + (uint64_t)(u >> 2) * db, idesc, 1u);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase); phase ^= 1;
      total += clock64() - t0;
    }
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = total;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}
}  // namespace yy

extern "C" int yy_umma_rate(int N, int layout_type, int lbo_a, int sbo_a, int lbo_b, int sbo_b, int a_step, int b_step,
                            int per_round, int iters, int n_ctas, long long* out_cycles_dev, void* stream) {
  using namespace yy;
  if (yy_device_count() == 0) return set_error(YY_ERR_NO_DEVICE, "no CUDA device");
  const int smem = 200 * 1024;
  YY_CUDA_OK(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_rate_kernel<<<n_ctas, 128, smem, (cudaStream_t)stream>>>(N, layout_type, lbo_a, sbo_a, lbo_b, sbo_b, a_step, b_step,
                                                                 per_round, iters, out_cycles_dev);
  YY_LAUNCH_CHECK();
  return YY_OK;
}

// ------------------------------------------------------------------------------------------------ L2 stream bench
// Developer tool: every CTA streams `total_bytes` from a shared `src_bytes` buffer (wrapping) through a ring of
// `slots` x `chunk` bytes with cp.async.bulk, no compute.  Reports SM cycles per CTA -> achievable L2->smem rate
// when all SMs pull the same weight image (what the tower kernel's producer does).
namespace yy {
__global__ void __launch_bounds__(256) l2_stream_kernel(const uint8_t* __restrict__ src, long long src_bytes, long long total_bytes,
                                                       int chunk, int slots, int warps, long long* out_cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[8][16];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int w = 0; w < 8; ++w) for (int s = 0; s < slots; ++s) mbar_init(smem_u32(&full[w][s]), 1); fence_barrier_init(); }
  __syncthreads();
  if (warp < warps) {   // each issuing warp streams its own share through its own sub-ring
    const long long n = total_bytes / chunk / warps;
    uint8_t* ring = smem + (size_t)warp * slots * chunk;
    const long long t0 = clock64();
    for (long long i = 0; i < n + slots; ++i) {
      if (i >= slots) mbar_wait(smem_u32(&full[warp][(i - slots) % slots]), (uint32_t)(((i - slots) / slots) & 1));
      if (i < n && elect_one()) {
        const int s = (int)(i % slots);
        mbar_arrive_expect_tx(smem_u32(&full[warp][s]), (uint32_t)chunk);
        bulk_g2s(smem_u32(ring + (size_t)s * chunk), src + ((i * warps + warp) * chunk) % src_bytes, (uint32_t)chunk, smem_u32(&full[warp][s]));
      }
      __syncwarp();
    }
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = clock64() - t0;
  }
}
}  // namespace yy

extern "C" int yy_l2_stream(const void* src_dev, long long src_bytes, long long total_bytes, int chunk, int slots, int warps, int n_ctas,
                            long long* out_cycles_dev, void* stream) {
  using namespace yy;
  if (yy_device_count() == 0) return set_error(YY_ERR_NO_DEVICE, "no CUDA device");
  if (slots < 1 || slots > 16 || warps < 1 || warps > 8 || chunk % 16 || (long long)chunk * slots * warps > 200 * 1024) return set_error(YY_ERR_INVALID, "bad ring");
  const int smem = chunk * slots * warps;
  YY_CUDA_OK(cudaFuncSetAttribute(l2_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  l2_stream_kernel<<<n_ctas, 256, smem, (cudaStream_t)stream>>>((const uint8_t*)src_dev, src_bytes, total_bytes, chunk, slots, warps, out_cycles_dev);
  YY_LAUNCH_CHECK();
  return YY_OK;
}

// ------------------------------------------------------------------------------------------------ tf32 MN-major probe
// Developer tool: C[128,N] = A * B^T with kind::tf32 where A is given TRANSPOSED in memory (At [K][128], "MN-major") and staged
// in shared memory in the canonical MN-major layout of the chosen swizzle mode; B [N][K] is K-major, no swizzle (known good).
//   variant 0: no swizzle       [k/8][m/4][8 k][16 B]                    SBO = 128 (m groups), LBO = 4096 (k groups)
//   variant 1: SWIZZLE_128B     [k/8][m/32][8 k][128 B], 16-byte chunk c of k-row r stored at chunk c ^ r; LBO = 1024 (m groups of 32),
//                               SBO = 4096 (k groups)
//   variant 2: the layout of variant 1 with layout type SWIZZLE_128B_BASE32B (rejected by the entry point: illegal memory access)
// Result on B200 (round 1): variants 0 and 1 complete and return C == 0 for every K and N tried, i.e. with these layouts the tf32
// MMA does not read a transposed A; the learner therefore stages transposed copies (yy_learn.cu).
namespace yy {
__global__ void __launch_bounds__(128) probe_tf32_mn_kernel(const float* __restrict__ At, const float* __restrict__ B, float* __restrict__ C,
                                                           int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* a_s = smem;                                   // K/8 groups x 4096 B
  uint8_t* b_s = smem + (size_t)(K / 8) * 4096;          // K-major: [K/4 planes][N rows][16 B]
  for (int i = threadIdx.x; i < K * 32; i += blockDim.x) {   // chunk (k, m4)
    const int k = i / 32, m4 = i % 32;
    const float4 v = *reinterpret_cast<const float4*>(At + (size_t)k * 128 + m4 * 4);
    const int kg = k >> 3, kr = k & 7;
    size_t off;
    if (variant == 0) off = (size_t)kg * 4096 + m4 * 128 + kr * 16;
    else { const int ng = m4 >> 3, c = m4 & 7; off = (size_t)kg * 4096 + ng * 1024 + kr * 128 + ((c ^ kr) * 16); }
    *reinterpret_cast<float4*>(a_s + off) = v;
  }
  for (int i = threadIdx.x; i < (K / 4) * N; i += blockDim.x) {
    const int kc = i / N, r = i % N;
    *reinterpret_cast<float4*>(b_s + (size_t)i * 16) = *reinterpret_cast<const float4*>(B + (size_t)r * K + kc * 4);
  }
  uint32_t ncols = 32; while ((int)ncols < N) ncols <<= 1;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), ncols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t lt = variant == 0 ? 0ull : (variant == 1 ? 2ull : 1ull);
    for (int k8 = 0; k8 < K / 8; ++k8) {
      const uint32_t a_addr = smem_u32(a_s) + (uint32_t)(k8 * 4096);
      const uint32_t b_addr = smem_u32(b_s) + (uint32_t)(2 * k8 * N * 16);
      const uint64_t ad = (variant == 0 ? smem_desc(a_addr, 4096, 128) : smem_desc(a_addr, 1024, 4096)) | (lt << 61);
      const uint64_t bd = smem_desc(b_addr, N * 16, 128);
      asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                   ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"(k8 > 0 ? 1u : 0u) : "memory");
    }
    tc_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
    tc_wait_ld();
    for (int j = 0; j < 16; ++j) C[(size_t)(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_base, ncols);
}
}  // namespace yy

extern "C" int yy_probe_tf32_mn(const float* at_dev, const float* b_dev, float* c_dev, int N, int K, int variant, void* stream) {
  using namespace yy;
  if (yy_device_count() == 0) return set_error(YY_ERR_NO_DEVICE, "no CUDA device");
  if (N < 16 || N > 256 || N % 16 || K < 8 || K % 8 || K > 128) return set_error(YY_ERR_INVALID, "probe: N in [16,256] step 16, K multiple of 8 up to 128");
  if (variant != 0 && variant != 1) return set_error(YY_ERR_INVALID, "probe: variant 0 or 1 (variant 2 raised an illegal memory access on sm_100a)");
  const size_t smem = (size_t)(K / 8) * 4096 + (size_t)(K / 4) * N * 16 + 1024;
  YY_CUDA_OK(cudaFuncSetAttribute(probe_tf32_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_tf32_mn_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(at_dev, b_dev, c_dev, N, K, variant);
  YY_LAUNCH_CHECK();
  return YY_OK;
}
