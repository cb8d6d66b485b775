// yy_gemm.cu -- tcgen05 GEMM building block: C[M,N] = act(A[M,K] * B[N,K]^T + bias), bf16 in, fp32 accumulate
// in TMEM.  Used for the policy / value fully-connected heads (neural_network.py:113-121) and, through
// yy_probe_umma, as the self-test that pins the shared-memory descriptor conventions of the tower kernel.
//
// Operands are plain row-major in global memory; cp.async (16 B = 8 bf16 along K) scatters them into the
// no-swizzle K-major core-matrix layout [k/8][row][8] in shared memory, which is exactly what the UMMA
// descriptor (yy_ptx.cuh: smem_desc) describes with LBO = rows*16 and SBO = 128.
#include <cuda_bf16.h>

#include "yy_common.cuh"
#include "yy_gemm.cuh"
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

// ------------------------------------------------------------------------------------------------ tiled GEMM
// grid = (ceil(M/128), N/BN).  128 threads.  BK = 64 per stage, 3-stage cp.async ring; thread 0 issues the
// MMAs and frees a stage with tcgen05.commit; all four warps drain TMEM in the epilogue.
constexpr int kGemmBK = 64;
constexpr int kGemmStages = 4;   // 96 KB: two CTAs per SM, so the 160 head tiles of a 4,096-board batch are one wave

struct GemmPair { GemmArgs p[2]; int ytiles0; };

template <int BN>
__device__ __forceinline__ void gemm_tile(const GemmArgs& g, int tile_m, int tile_n);

template <int BN>
__global__ void __launch_bounds__(128) gemm_bf16_tn_kernel(GemmArgs g) { gemm_tile<BN>(g, blockIdx.x, blockIdx.y); }

template <int BN>
__global__ void __launch_bounds__(128) gemm_bf16_tn_pair_kernel(GemmPair gp) {
  if ((int)blockIdx.y < gp.ytiles0) gemm_tile<BN>(gp.p[0], blockIdx.x, blockIdx.y);
  else gemm_tile<BN>(gp.p[1], blockIdx.x, blockIdx.y - gp.ytiles0);
}

template <int BN>
__device__ __forceinline__ void gemm_tile(const GemmArgs& g, int tile_m, int tile_n) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mma_done[kGemmStages];
  __shared__ __align__(8) uint64_t acc_done;
  __shared__ uint32_t tmem_base_s;
  // k-chunk planes are padded by one 16-byte row (LBO = (rows+1)*16) so that the 8 chunks of one global 128-byte
  // line, fetched by 8 consecutive threads, land in 8 different bank groups
  constexpr int A_LBO = (128 + 1) * 16, B_LBO = (BN + 1) * 16;
  constexpr int A_STAGE = (kGemmBK / 8) * A_LBO;
  constexpr int B_STAGE = (kGemmBK / 8) * B_LBO;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + kGemmStages * A_STAGE;
  const int m0 = tile_m * 128, n0 = tile_n * BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t NCOLS = BN < 32 ? 32 : BN;
  if (tid == 0) {
    for (int s = 0; s < kGemmStages; ++s) mbar_init(smem_u32(&mma_done[s]), 1);
    mbar_init(smem_u32(&acc_done), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), NCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int nkb = (g.K + kGemmBK - 1) / kGemmBK;

  auto load_stage = [&](int kb, int s) {
    const int k0 = kb * kGemmBK;
    // 8 consecutive threads fetch the 8 chunks (one contiguous 128-byte line) of one row
    const int kc = tid & 7;
    const bool kv = (k0 + kc * 8) < g.K;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = (tid >> 3) + 16 * i;
      const bool rv = (m0 + r) < g.M;
      cp_async16(smem_u32(a_s + s * A_STAGE + kc * A_LBO + r * 16), g.A + (size_t)(rv ? (m0 + r) : 0) * g.lda + k0 + kc * 8, rv && kv);
    }
#pragma unroll
    for (int i = 0; i < BN / 16; ++i) {
      const int r = (tid >> 3) + 16 * i;
      const bool rv = (n0 + r) < g.N;
      cp_async16(smem_u32(b_s + s * B_STAGE + kc * B_LBO + r * 16), g.B + (size_t)(rv ? (n0 + r) : 0) * g.ldb + k0 + kc * 8, rv && kv);
    }
    cp_async_commit();
  };

  // prologue
  for (int p = 0; p < kGemmStages - 1; ++p) { if (p < nkb) load_stage(p, p); else cp_async_commit(); }
  uint32_t done_phase_bits = 0;   // bit s = parity to wait for on mma_done[s]
  const uint32_t idesc = idesc_bf16(128, BN);
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % kGemmStages;
    // issue the load for k-block kb + stages - 1 into the stage that k-block kb-1 used
    {
      const int nk = kb + kGemmStages - 1, ns = nk % kGemmStages;
      if (nk < nkb) {
        if (kb >= 1) { mbar_wait(smem_u32(&mma_done[ns]), (done_phase_bits >> ns) & 1u); done_phase_bits ^= 1u << ns; }
        load_stage(nk, ns);
      } else {
        cp_async_commit();
      }
    }
    cp_async_wait<kGemmStages - 1>();
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k16 = 0; k16 < kGemmBK / 16; ++k16) {
        uint64_t ad = smem_desc(smem_u32(a_s + s * A_STAGE) + 2 * k16 * A_LBO, A_LBO, 128);
        uint64_t bd = smem_desc(smem_u32(b_s + s * B_STAGE) + 2 * k16 * B_LBO, B_LBO, 128);
        tc_mma_bf16(tmem_base, ad, bd, idesc, (kb > 0 || k16 > 0) ? 1u : 0u);
      }
      tc_commit(smem_u32(&mma_done[s]));
      if (kb == nkb - 1) tc_commit(smem_u32(&acc_done));
    }
  }
  mbar_wait(smem_u32(&acc_done), 0);
  tc_fence_after();
  const int row = m0 + warp * 32 + lane;
  for (int c = 0; c < BN; c += 16) {
    uint32_t r[16];
    tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
    tc_wait_ld();
    if (row < g.M) {
      float out[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float v = __uint_as_float(r[j]);
        if (g.bias) v += g.bias[n0 + c + j];
        if (g.relu) v = fmaxf(v, 0.0f);
        out[j] = v;
      }
      float4* dst = reinterpret_cast<float4*>(g.C + (size_t)row * g.ldc + n0 + c);
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, NCOLS);
}

template <int BN>
static int launch_gemm(const GemmArgs& g, cudaStream_t s) {
  constexpr int smem = kGemmStages * (kGemmBK / 8) * ((128 + 1) * 16 + (BN + 1) * 16);
  static bool attr_done = false;
  if (!attr_done) {
    YY_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  dim3 grid((unsigned)((g.M + 127) / 128), (unsigned)(g.N / BN));
  gemm_bf16_tn_kernel<BN><<<grid, 128, smem, s>>>(g);
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int gemm_bf16_tn_pair(const GemmArgs& p0, const GemmArgs& p1, cudaStream_t s) {
  if (p0.M <= 0) return YY_OK;
  if (p0.M != p1.M || p0.N % 64 || p1.N % 64 || p0.K % 8 || p1.K % 8 || p0.lda % 8 || p1.lda % 8 || p0.ldb % 8 || p1.ldb % 8)
    return set_error(YY_ERR_INVALID, "gemm pair: equal M, N multiples of 64, K/lda/ldb multiples of 8 required");
  constexpr int BN = 64;
  constexpr int smem = kGemmStages * (kGemmBK / 8) * ((128 + 1) * 16 + (BN + 1) * 16);
  static bool attr_done = false;
  if (!attr_done) {
    YY_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  GemmPair gp; gp.p[0] = p0; gp.p[1] = p1; gp.ytiles0 = p0.N / BN;
  dim3 grid((unsigned)((p0.M + 127) / 128), (unsigned)(p0.N / BN + p1.N / BN));
  gemm_bf16_tn_pair_kernel<BN><<<grid, 128, smem, s>>>(gp);
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int gemm_bf16_tn(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0) return YY_OK;
  if (g.K % 8 || g.lda % 8 || g.ldb % 8 || g.ldc % 4) return set_error(YY_ERR_INVALID, "gemm: K, lda, ldb must be multiples of 8, ldc of 4");
  if (g.N % 64 == 0) return launch_gemm<64>(g, s);
  if (g.N % 32 == 0) return launch_gemm<32>(g, s);
  if (g.N % 16 == 0) return launch_gemm<16>(g, s);
  return set_error(YY_ERR_INVALID, "gemm: N must be a multiple of 16 (got %d)", g.N);
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
