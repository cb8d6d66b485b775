// yy_learn.cu -- learner-step primitives: the kernels behind one optimisation step of
//   AlphaZeroTrainer.train            src/yin_yang/ai/trainer.py:67-161  (Adam lr 1e-3, weight_decay 1e-4,
//                                     CrossEntropyLoss(soft targets) + MSELoss, batch 64, nnet.train())
//   YinYangNeuralNetwork.forward      src/yin_yang/ai/neural_network.py:94-123 in TRAINING mode (batch-norm batch statistics)
// Activations live in HBM as fp32 [positions][channels] (positions = batch x cells, channel contiguous).  Every
// convolution / linear layer and both of its gradients are ONE GEMM kernel, C[M,N] = A[M,K] * B[N,K]^T on the 5th-gen
// tensor cores (tcgen05.mma kind::tf32, fp32 accumulation in TMEM; 3xTF32 by default = fp32-level results).  The im2col
// of a 3x3 convolution is implicit in the A-operand loader of the forward and backward-data passes (a tap is a row
// offset, borders are zero-filled by cp.async); the weight gradient, whose reduction index is the position, reads
// transposed copies ([channels][positions]) written by two small kernels.  (tcgen05 can read a transposed, "MN-major",
// operand itself, but with kind::tf32 and the no-swizzle layout the MMA returned zeros in every descriptor variant
// tried here; the 32-bit case seems to need the 128B/32B-atom swizzle mode.)  Split-K partial tiles go to a workspace
// and a reducer adds them in a fixed order (bias, skip share, ReLU fused there): every result is bit-reproducible.
// Batch-norm statistics are float64.  A batch of 64 boards is 4,096 positions: the step is a few hundred
// microsecond-sized kernels in one CUDA graph (learner.py).
#include "yy_common.cuh"
#include "yy_ptx.cuh"
#include <stdlib.h>

namespace yy {
using namespace ptx;

// Programmatic dependent launch: every kernel of the learner step is launched with programmatic stream serialisation, lets the
// next kernel of the stream start at once (its blocks become resident as SMs free up and run their prologue) and waits for
// its predecessors' memory before its first global access.  A step is ~170 dependent microsecond kernels: what PDL hides is
// the launch latency between them.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// D[tmem] (+)= A[smem] * B[smem]^T, tf32 x tf32 -> fp32 (K = 8 per instruction), issued by ONE thread
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Instruction descriptor for kind::tf32: D fp32, A/B tf32 (format 2), both K-major, dense, MxN tile.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ GEMM
// One CTA per (128-row, tile_n-column, K-slice) tile; stages of 32 K-values; operands go global -> shared with 16-byte
// loads through registers (LDG.128 -> hi/lo split -> STS.128; the 16-byte cp.async scatter measured one chunk per cycle per
// SM and bounded the kernel) into UMMA's no-swizzle K-major core-matrix layout: [k/4 planes][row][16 B], plane pitch
// rows*16+16 B (LBO = pitch, SBO = 128; the 16 B of padding keep the stores bank-conflict free); zeros past the edges.
// X3 (default precision of the learner): every operand chunk is split in registers on its way to shared memory
// into hi = the 19 bits a TF32 multiplier sees and lo = x - hi, and each K-slice runs three MMAs (lo*hi + hi*lo + hi*hi)
// -- "3xTF32": products carry ~21 mantissa bits, i.e. fp32-level results (the reference trains in fp32; with plain TF32
// ~4e-4 of the ReLU masks flip and the gradients differ from fp32 by 5-9 % in L2).
constexpr int kGemmKStage = 32;                     // K-values per stage = 4 K-slices of 8
constexpr int kPlaneA = 128 * 16 + 16;
constexpr int kRegionA = 8 * kPlaneA;               // bytes of a stage's A tile
enum { OP_K = YY_OP_K, OP_K_CONV = YY_OP_K_CONV, OP_K_CONVT = YY_OP_K_CONVT };

struct GemmArgs {
  const float* A; const float* B; float* C; const float* bias; float* ws;
  const uint8_t* Bpack;                             // B pre-split (hi / lo) and pre-tiled per K stage by yy_lrn_pack_b (3xTF32, N = tile_n = 128)
  int lda, ldb, ldc, M, N, K, tile_n, k_per_split, relu, accumulate, a_mode, b_mode;
  int rows, cols, cin, flip;                        // convolution geometry of the implicit A operand (cin: gathered channels)
  int cz;                                           // CTAs per cluster along z: the K slices whose partial tiles are summed on chip
  double* bn_sums;                                  // fused batch-norm statistics when the cluster produces the final C
  const float* st_out; const float* st_y; const float* st_mi; int st_ldo, st_ldy;   // st_out != NULL: backward statistics of C (a gradient)
  long long* dbg;                                   // developer tool: clock64 stamps of CTA 0 (yy_lrn_gemm_debug_stamps)
};

// is the (dx,dy) neighbour of the cell of position p on the board?  offset = dx*cols + dy
__device__ __forceinline__ bool tap_ok(int p, int tap, int rows, int cols, int flip, int* offset) {
  const int cell = p % (rows * cols), x = cell / cols, y = cell - x * cols;
  int dx = tap / 3 - 1, dy = tap - (tap / 3) * 3 - 1;
  if (flip) { dx = -dx; dy = -dy; }
  *offset = dx * cols + dy;
  return (unsigned)(x + dx) < (unsigned)rows && (unsigned)(y + dy) < (unsigned)cols;
}

#define YY_STAMP(i) do { if (stamp && (i) < 120) g.dbg[(i)] = clock64(); } while (0)
// Warp roles: warps 0-7 = producers (two warps per scheduler; the loads of stage kt+1 are in flight while stage kt is
// split and stored) and, at the end, the epilogue (warps w and w+4 share TMEM lanes
// 32(w&3).. and alternate 16-column chunks); warp 8 = MMA issuer.  full[s] (256 arrivals: every producer thread after its
// chunks of stage s are in place and fenced for the async proxy), free[s] (tcgen05.commit: the MMAs that read stage s
// have completed), done (accumulator complete).  All per-row address arithmetic (board coordinates of the implicit
// im2col included) is done once per CTA.
template <int S, bool X3>
__global__ void __launch_bounds__(288) gemm_tf32_kernel(GemmArgs g) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[S];
  __shared__ __align__(8) uint64_t free_bar[S];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool stamp = g.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
  YY_STAMP(100);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * g.tile_n;
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const int KT = (k_end - k_begin + kGemmKStage - 1) / kGemmKStage;
  const int planeA = kPlaneA, planeB = g.tile_n * 16 + 16;
  const int half_bytes = kRegionA + 8 * planeB;               // one copy (hi) of a stage's A and B tiles
  const int stage_bytes = X3 ? 2 * half_bytes : half_bytes;
  uint32_t ncols = 32; while ((int)ncols < g.tile_n) ncols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(smem_u32(&full_bar[s]), g.Bpack ? 257 : 256); mbar_init(smem_u32(&free_bar[s]), 1); }
    mbar_init(smem_u32(&done_bar), 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&tmem_base_s), ncols);
  // OP_K_CONVT: per position of this CTA's K range, which neighbours of its cell are on the board (same 4 bits as below)
  uint8_t* edge_s = smem + (size_t)S * stage_bytes;
  if (g.b_mode == OP_K_CONVT)
    for (int i = tid; i < k_end - k_begin; i += blockDim.x) {
      const int cell = (k_begin + i) % (g.rows * g.cols), x = cell / g.cols, y = cell - x * g.cols;
      edge_s[i] = (uint8_t)((x > 0 ? 1u : 0u) | (x < g.rows - 1 ? 2u : 0u) | (y > 0 ? 4u : 0u) | (y < g.cols - 1 ? 8u : 0u));
    }
  YY_STAMP(0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t smem0 = smem_u32(smem);
  YY_STAMP(1);

  if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = idesc_tf32(128, g.tile_n);
    const uint64_t a0 = smem_desc(smem0, planeA, 128), b0 = smem_desc(smem0 + kRegionA, planeB, 128);
    const uint64_t a_step = (uint64_t)((2 * planeA) >> 4), b_step = (uint64_t)((2 * planeB) >> 4), lo = (uint64_t)(half_bytes >> 4);
    for (int kt = 0; kt < KT; ++kt) {
      const int s = kt % S;
      mbar_wait(smem_u32(&full_bar[s]), (uint32_t)((kt / S) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint64_t st = (uint64_t)((s * stage_bytes) >> 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t ad = a0 + st + j * a_step, bd = b0 + st + j * b_step;
          const uint32_t first = (kt > 0 || j > 0) ? 1u : 0u;
          if (X3) {
            tc_mma_tf32(tmem_base, ad + lo, bd, idesc, first);
            tc_mma_tf32(tmem_base, ad, bd + lo, idesc, 1u);
            tc_mma_tf32(tmem_base, ad, bd, idesc, 1u);
          } else {
            tc_mma_tf32(tmem_base, ad, bd, idesc, first);
          }
        }
        tc_commit(smem_u32(&free_bar[s]));
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(smem_u32(&done_bar));
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ producers
    // Chunk ownership: lane = (k-chunk of the stage, 0..7) + 8 * (row within a 4-row group): one LDG.128 of a warp reads
    // 128 contiguous bytes of 4 rows (4 cache lines; 64 bytes of 8 rows cost twice the L1 tag lookups) and its STS.128
    // fills 64 bytes in each of the 8 chunk planes (conflict free: plane pitch = 16 mod 128).  Warp w owns A rows
    // 16w..16w+15 (4 chunks per thread) and B rows 4(w + 8j) + r4.
    const int c8 = lane & 7, r4 = lane >> 3, wq = warp & 3, wh = warp >> 2;
    const bool convA = g.a_mode == OP_K_CONV;
    const float* arow0 = g.A + (size_t)(m0 + 16 * warp + r4) * g.lda;     // row group gi adds 4*gi rows
    const float* brow0 = g.B + (size_t)(n0 + 4 * warp + r4) * g.ldb;      // row group j adds 32*j rows
    // per owned A row (gi = 0..3): bit 4gi+0/1/2/3 = the neighbour above / below / left / right of its cell is on the board
    uint32_t edge = 0, rowok = 0;
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) {
      const int p = m0 + 16 * warp + 4 * gi + r4;
      if (p < g.M) {
        rowok |= 1u << gi;
        if (convA) {
          const int cell = p % (g.rows * g.cols), x = cell / g.cols, y = cell - x * g.cols;
          edge |= ((x > 0 ? 1u : 0u) | (x < g.rows - 1 ? 2u : 0u) | (y > 0 ? 4u : 0u) | (y < g.cols - 1 ? 8u : 0u)) << (4 * gi);
        }
      }
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // OP_K_CONVT: B row n = tap*cin + ci is row ci of the transposed activation shifted by d(tap) positions
    const bool convB = g.b_mode == OP_K_CONVT;
    const bool packedB = X3 && g.Bpack != nullptr;
    // packed B: one thread streams the stage's hi and lo tiles (2 x 16.5 KB, already in the core-matrix layout) with two bulk
    // async copies that complete on the stage's full barrier -- no LSU requests, no split work for the weights
    auto bulk_b = [&](int kt) {
      const int s = kt % S;
      const uint32_t bar = smem_u32(&full_bar[s]);
      const uint32_t dst = smem0 + (uint32_t)(s * stage_bytes + kRegionA);
      const uint8_t* src = g.Bpack + (size_t)(k_begin / kGemmKStage + kt) * (2 * kRegionA);
      mbar_arrive_expect_tx(bar, 2u * kRegionA);
      bulk_g2s(dst, src, kRegionA, bar);
      bulk_g2s(dst + (uint32_t)half_bytes, src + kRegionA, kRegionA, bar);
    };
    const float* bptr[4] = {g.B, g.B, g.B, g.B};
    uint32_t bneed = 0, bok = 0;                              // bok bit j: B row 4(w+8j)+r4 is inside the tile and inside N
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = 4 * (warp + 8 * j) + r4, n = n0 + row;
      if (row < g.tile_n && n < g.N) {
        bok |= 1u << j;
        if (convB) {
          const int tap = n / g.cin, ci = n - tap * g.cin, dx = tap / 3 - 1, dy = tap - (tap / 3) * 3 - 1;
          bneed |= ((dx < 0 ? 1u : 0u) | (dx > 0 ? 2u : 0u) | (dy < 0 ? 4u : 0u) | (dy > 0 ? 8u : 0u)) << (4 * j);
          bptr[j] = g.B + (long long)ci * g.ldb + (dx * g.cols + dy);
        }
      }
    }
    // global -> registers (read-only path), zeros past the edges / outside the board
    auto load_regs = [&](float4 (&ra)[4], float4 (&rb)[4], int kt) {
      const int k = k_begin + kt * kGemmKStage + c8 * 4;      // this thread's k-chunk of every row it owns
      const bool kok = k < k_end;
      {
        long long aoff = k;                                   // element offset added to a row's base
        uint32_t need = 0;                                    // edge bits this tap needs
        if (convA) {
          const int tap = k / g.cin, ci = k - tap * g.cin;
          int dx = tap / 3 - 1, dy = tap - (tap / 3) * 3 - 1;
          if (g.flip) { dx = -dx; dy = -dy; }
          need = (dx < 0 ? 1u : 0u) | (dx > 0 ? 2u : 0u) | (dy < 0 ? 4u : 0u) | (dy > 0 ? 8u : 0u);
          aoff = (long long)(dx * g.cols + dy) * g.lda + ci;
        }
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          const bool ok = kok && ((rowok >> gi) & 1u) && ((need & ~(edge >> (4 * gi))) & 15u) == 0;
          ra[gi] = ok ? __ldg(reinterpret_cast<const float4*>(arow0 + (size_t)(4 * gi) * g.lda + aoff)) : zero4;
        }
      }
      if (packedB) {
        // B arrives by bulk copy (below)
      } else if (convB) {
        // (two aligned LDG.128 + selects instead of these 4 LDG.32 per chunk measured slower: 25 -> 32 us per weight-gradient GEMM)
        uint32_t e4 = 0;                                      // edge bits of the chunk's 4 positions
        if (kok) e4 = *reinterpret_cast<const uint32_t*>(edge_s + (k - k_begin));   // K is a multiple of 4: whole chunks
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool ok = ((bok >> j) & 1u) && (k + e) < k_end && (((bneed >> (4 * j)) & ~(e4 >> (8 * e))) & 15u) == 0;
            v[e] = ok ? __ldg(bptr[j] + k + e) : 0.f;
          }
          rb[j] = make_float4(v[0], v[1], v[2], v[3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = kok && ((bok >> j) & 1u);
          rb[j] = ok ? __ldg(reinterpret_cast<const float4*>(brow0 + (size_t)(32 * j) * g.ldb + k)) : zero4;
        }
      }
    };
    // registers -> the stage's core-matrix layout (hi / lo copies for 3xTF32), then publish the stage
    auto put = [&](uint8_t* dst, const float4& v) {
      if (X3) {
        const float4 h = make_float4(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                                     __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
        *reinterpret_cast<float4*>(dst) = h;
        *reinterpret_cast<float4*>(dst + half_bytes) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      } else {
        *reinterpret_cast<float4*>(dst) = v;
      }
    };
    auto store_stage = [&](const float4 (&ra)[4], const float4 (&rb)[4], int s) {
      uint8_t* st = smem + (size_t)s * stage_bytes;
      uint8_t* a = st + c8 * planeA + (16 * warp + r4) * 16;
#pragma unroll
      for (int gi = 0; gi < 4; ++gi) put(a + gi * 64, ra[gi]);
      uint8_t* b = st + kRegionA + c8 * planeB + (4 * warp + r4) * 16;
      if (!packedB) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (4 * (warp + 8 * j) + r4 < g.tile_n) put(b + j * 512, rb[j]);
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&full_bar[s]));
    };
    // (Tried: three rotating register sets = two stages of loads in flight: 168 registers, spills in the 3xTF32 variant and
    // no gain -- the K-iteration takes ~3,000 cycles whatever the number of CTAs (32..288, tools/gemm_phases.py): it is
    // bounded by the L1 request path of the 64 LDG.128 per stage, 8 cache lines each, not by bytes in flight or by L2.)
    // Everything above (barriers, TMEM, the board coordinates of this thread's rows: ~2,500 cycles of integer division) ran
    // under the previous kernel's tail; its memory is needed from here on.  Only these warps touch global memory.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    YY_STAMP(103);
    float4 ra[4], rb[4], na[4], nb[4];
    if (KT > 0) load_regs(ra, rb, 0);
    if (packedB && tid == 0 && KT > 0) bulk_b(0);
    for (int kt = 0; kt < KT; ++kt) {
      const int s = kt % S;
      YY_STAMP(4 + 6 * kt);
      if (kt + 1 < KT) load_regs(na, nb, kt + 1);             // next stage's loads in flight while this one is stored
      YY_STAMP(5 + 6 * kt);
      if (kt + 1 < KT) {                                      // the slot of stage kt+1: the MMAs that read it S iterations earlier
        if (kt + 1 >= S) {
          if (lane == 0) mbar_wait(smem_u32(&free_bar[(kt + 1) % S]), (uint32_t)((((kt + 1) / S) - 1) & 1));
          __syncwarp();
        }
        if (packedB && tid == 0) bulk_b(kt + 1);
      }
      YY_STAMP(6 + 6 * kt);
      store_stage(ra, rb, s);
      YY_STAMP(7 + 6 * kt);
#pragma unroll
      for (int i = 0; i < 4; ++i) { ra[i] = na[i]; rb[i] = nb[i]; }
    }
    YY_STAMP(2);
    if (lane == 0) mbar_wait(smem_u32(&done_bar), 0);
    __syncwarp();
    tc_fence_after();
    YY_STAMP(119);

    // epilogue: warps w and w+4 read TMEM lanes 32(w&3).. (alternate 16-column chunks) and park the tile in shared memory
    // (the stages are free now; row pitch tile_n+4 floats keeps both sides conflict free); then every warp writes whole
    // rows, 512 contiguous bytes per instruction.  With split-K the tile goes to the workspace ([slice][M][N]) and
    // gemm_reduce_kernel finishes it; otherwise bias / skip share / ReLU are applied here.
    float* tile = reinterpret_cast<float*>(smem);
    const int pitch = g.tile_n + 4;
    if (KT > 0) {
      for (int c = 16 * wh; c < g.tile_n; c += 32) {
        uint32_t r[16];
        tc_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c, r);
        tc_wait_ld();
        float* dst = tile + (wq * 32 + lane) * pitch + c;
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4)
          *reinterpret_cast<float4*>(dst + j4) = make_float4(__uint_as_float(r[j4]), __uint_as_float(r[j4 + 1]), __uint_as_float(r[j4 + 2]), __uint_as_float(r[j4 + 3]));
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");            // the 8 producer / epilogue warps
    const int n4 = g.tile_n >> 2;                               // float4 columns of the tile
    if (g.cz == 1) {
      const bool partial = gridDim.z > 1;
      for (int idx = tid; idx < 128 * n4; idx += 256) {
        const int rr = idx / n4, cc = (idx - rr * n4) * 4, row = m0 + rr, n = n0 + cc;
        if (row < g.M && n < g.N) {                            // N is a multiple of 4
          float4 v = KT > 0 ? *reinterpret_cast<const float4*>(tile + rr * pitch + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
          float* crow = partial ? g.ws + ((size_t)blockIdx.z * g.M + row) * g.N : g.C + (size_t)row * g.ldc;
          if (!partial) {
            if (g.bias) { v.x += g.bias[n]; v.y += g.bias[n + 1]; v.z += g.bias[n + 2]; v.w += g.bias[n + 3]; }
            if (g.accumulate) { const float4 o = *reinterpret_cast<const float4*>(crow + n); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            if (g.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          }
          *reinterpret_cast<float4*>(crow + n) = v;
        }
      }
    }
    YY_STAMP(3);
  }
  if (g.cz > 1) {
    // Split-K inside a thread-block cluster: the cz CTAs of a cluster hold the partial tiles of cz consecutive K slices in
    // their shared memory; CTA r finishes rows [r*128/cz, (r+1)*128/cz) of the tile by reading all cz partials through
    // distributed shared memory in rank order (deterministic), so no workspace round trip and no reducer launch.  With
    // more slices than a cluster holds (gridDim.z > cz) the cluster's sum goes to the workspace as ONE slice.
    cluster_sync_all();
    if (warp < 8) {
      const float* tile = reinterpret_cast<const float*>(smem);
      const uint32_t tile_s = smem_u32(smem);
      const int pitch = g.tile_n + 4, n4 = g.tile_n >> 2, rows_per = 128 / g.cz;
      const uint32_t rank = cluster_ctarank();
      const int groups = gridDim.z / g.cz, grp = blockIdx.z / g.cz;
      const bool partial = groups > 1;
      __shared__ float red[2][256][4];
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
      // The partial tiles of the other CTAs come through distributed shared memory, ~20 B per cycle and SM whatever the number of
      // loads in flight (measured with 8 per thread: no change): 48 KB of remote partials per CTA make this phase ~5,000 of the
      // kernel's ~25,000 cycles (tools/gemm_phases.py) -- still cheaper than the workspace round trip + reducer launch it replaced.
      for (int idx = tid; idx < rows_per * n4; idx += 256) {
        const int rr = (int)rank * rows_per + idx / n4, cc = (idx % n4) * 4, row = m0 + rr, n = n0 + cc;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t off = (uint32_t)((rr * pitch + cc) * 4);
        for (int j = 0; j < g.cz; ++j) {
          float4 pv;
          const uint32_t addr = mapa_u32(tile_s + off, (uint32_t)j);
          asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(pv.x), "=f"(pv.y), "=f"(pv.z), "=f"(pv.w) : "r"(addr) : "memory");
          v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w;
        }
        if (row < g.M && n < g.N) {
          float* crow = partial ? g.ws + ((size_t)grp * g.M + row) * g.N : g.C + (size_t)row * g.ldc;
          if (!partial) {
            if (g.bias) { v.x += g.bias[n]; v.y += g.bias[n + 1]; v.z += g.bias[n + 2]; v.w += g.bias[n + 3]; }
            if (g.accumulate) { const float4 o = *reinterpret_cast<const float4*>(crow + n); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            if (g.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            if (g.bn_sums) {
              if (g.st_out == nullptr) {                       // forward: sum, sum of squares of C
                s1[0] += v.x; s1[1] += v.y; s1[2] += v.z; s1[3] += v.w;
                s2[0] += v.x * v.x; s2[1] += v.y * v.y; s2[2] += v.z * v.z; s2[3] += v.w * v.w;
              } else {                                         // backward: C is dOut of a batch norm + ReLU: sum dZ*xhat, sum dZ
                const float4 o = *reinterpret_cast<const float4*>(g.st_out + (size_t)row * g.st_ldo + n);
                const float4 y = *reinterpret_cast<const float4*>(g.st_y + (size_t)row * g.st_ldy + n);
                const float4 mu = *reinterpret_cast<const float4*>(g.st_mi + n), is = *reinterpret_cast<const float4*>(g.st_mi + g.N + n);
                const float dz0 = o.x > 0.f ? v.x : 0.f, dz1 = o.y > 0.f ? v.y : 0.f, dz2 = o.z > 0.f ? v.z : 0.f, dz3 = o.w > 0.f ? v.w : 0.f;
                s1[0] += dz0 * ((y.x - mu.x) * is.x); s1[1] += dz1 * ((y.y - mu.y) * is.y);
                s1[2] += dz2 * ((y.z - mu.z) * is.z); s1[3] += dz3 * ((y.w - mu.w) * is.w);
                s2[0] += dz0; s2[1] += dz1; s2[2] += dz2; s2[3] += dz3;
              }
            }
          }
          *reinterpret_cast<float4*>(crow + n) = v;
        }
      }
      (void)tile;
      if (g.bn_sums && !partial) {                             // host guarantees 256 % n4 == 0: a thread keeps one column group
#pragma unroll
        for (int j = 0; j < 4; ++j) { red[0][tid][j] = s1[j]; red[1][tid][j] = s2[j]; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < n4) {
          for (int l = 1; l < 256 / n4; ++l)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s1[j] += red[0][l * n4 + tid][j]; s2[j] += red[1][l * n4 + tid][j]; }
          const int n = n0 + tid * 4;
          if (n < g.N)
#pragma unroll
            for (int j = 0; j < 4; ++j) { atomicAdd(&g.bn_sums[n + j], (double)s1[j]); atomicAdd(&g.bn_sums[g.N + n + j], (double)s2[j]); }
        }
      }
    }
    YY_STAMP(101);
    cluster_sync_all();                                        // nobody leaves while its tile is still being read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, ncols);
  YY_STAMP(102);
}

// C = [C +] bias + sum_z ws[z] [ReLU], slices added in index order (deterministic).  A block finishes 4 slabs of
// 256/N4 whole rows; bn_sums (optional, float64 [2N], zero on entry) receives the per-column sum / sum of squares of the
// finished C -- the statistics of the batch norm that follows a convolution, saving its own pass over C.
__global__ void __launch_bounds__(256) gemm_reduce_kernel(const float* __restrict__ ws, int splits, float* __restrict__ C, int ldc, int M, int N4,
                                                         const float* __restrict__ bias, int relu, int accumulate, double* __restrict__ bn_sums) {
  pdl_enter();
  __shared__ float red[2][256][4];
  const int cg = threadIdx.x % N4, n = cg * 4, rl = threadIdx.x / N4, lanes = 256 / N4;   // host guarantees 256 % N4 == 0 when bn_sums
  const size_t slice = (size_t)M * N4 * 4;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  float bi[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) { bi[0] = bias[n]; bi[1] = bias[n + 1]; bi[2] = bias[n + 2]; bi[3] = bias[n + 3]; }
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) {
    const int m = (blockIdx.x * 4 + sl) * lanes + rl;
    if (m < M && rl < lanes) {
      float a[4];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(ws + (size_t)m * N4 * 4 + n);
      for (int z = 1; z < splits; ++z) {
        const float4 b = *reinterpret_cast<const float4*>(ws + z * slice + (size_t)m * N4 * 4 + n);
        a[0] += b.x; a[1] += b.y; a[2] += b.z; a[3] += b.w;
      }
      float* c = C + (size_t)m * ldc + n;
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] += bi[j];
      if (accumulate) { const float4 o = *reinterpret_cast<const float4*>(c); a[0] += o.x; a[1] += o.y; a[2] += o.z; a[3] += o.w; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (relu) a[j] = fmaxf(a[j], 0.f);
        s1[j] += a[j]; s2[j] += a[j] * a[j];
      }
      *reinterpret_cast<float4*>(c) = *reinterpret_cast<float4*>(a);
    }
  }
  if (bn_sums) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[0][threadIdx.x][j] = s1[j]; red[1][threadIdx.x][j] = s2[j]; }
    __syncthreads();
    if (threadIdx.x < N4) {
      for (int l = 1; l < lanes; ++l)
#pragma unroll
        for (int j = 0; j < 4; ++j) { s1[j] += red[0][l * N4 + cg][j]; s2[j] += red[1][l * N4 + cg][j]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) { atomicAdd(&bn_sums[n + j], (double)s1[j]); atomicAdd(&bn_sums[4 * N4 + n + j], (double)s2[j]); }
    }
  }
}

// the same reduction for column counts that do not tile a 256-thread block
__global__ void __launch_bounds__(256) gemm_reduce_flat_kernel(const float* __restrict__ ws, int splits, float* __restrict__ C, int ldc, int M, int N4,
                                                              const float* __restrict__ bias, int relu, int accumulate) {
  pdl_enter();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * N4) return;
  const int m = (int)(idx / N4), n = (int)(idx % N4) * 4;
  const size_t slice = (size_t)M * N4 * 4;
  float4 a = *reinterpret_cast<const float4*>(ws + (size_t)m * N4 * 4 + n);
  for (int z = 1; z < splits; ++z) {
    const float4 b = *reinterpret_cast<const float4*>(ws + z * slice + (size_t)m * N4 * 4 + n);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  if (bias) { a.x += bias[n]; a.y += bias[n + 1]; a.z += bias[n + 2]; a.w += bias[n + 3]; }
  float* c = C + (size_t)m * ldc + n;
  if (accumulate) { const float4 o = *reinterpret_cast<const float4*>(c); a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
  if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
  *reinterpret_cast<float4*>(c) = a;
}

// ------------------------------------------------------------------------------------------------ layout / reduction kernels
// out[c][r] = in[r][c], for blockIdx.z = 0..batch-1 matrices at fixed strides (all layer inputs of the tower in one launch)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int ldo, int R, int C,
                                                       long long in_stride, long long out_stride) {
  pdl_enter();
  __shared__ float tile[32][33];
  in += (size_t)blockIdx.z * in_stride; out += (size_t)blockIdx.z * out_stride;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < R && c < C) ? in[(size_t)r * ldi + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < C && r < R) out[(size_t)c * ldo + r] = tile[tx][j];
  }
}
// colT[t*C + c][p] = X[p + d(t)][c] (zero outside the board): the transposed im2col the weight-gradient GEMM reads
// (its reduction index is the position).  One block per (128 positions, 32 channels, tap): all loads first, then the
// transposed stores (128 consecutive positions of one channel row per warp pass).
__global__ void __launch_bounds__(256) im2col_t_kernel(const float* __restrict__ X, int ldx, float* __restrict__ colT, int ldo, int P, int rows,
                                                      int cols, int C) {
  pdl_enter();
  __shared__ float tile[128][33];
  const int p0 = blockIdx.x * 128, c0 = blockIdx.y * 32, tap = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int p = p0 + ty + 8 * i, c = c0 + tx;
    v[i] = 0.f;
    if (p < P && c < C) {
      int off;
      if (tap_ok(p, tap, rows, cols, 0, &off)) v[i] = X[(size_t)(p + off) * ldx + c];
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) tile[ty + 8 * i][tx] = v[i];
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    if (c < C) {
      float* dst = colT + (size_t)(tap * C + c) * ldo + p0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pp = tx + 32 * i;
        if (p0 + pp < P) dst[pp] = tile[pp][j];
      }
    }
  }
}
// Wt[ci][t*Cout + co] = W[co][t*Cin + ci]  (B operand of the backward-data GEMM; the tap flip lives in the A gather)
// Layer l = blockIdx.y: W = params + offsets[l] -> Wt + l*Cout*9*Cin (all layers of the tower in one launch).
__global__ void __launch_bounds__(256) conv_weight_t_kernel(const float* __restrict__ params, const long long* __restrict__ offsets,
                                                           float* __restrict__ Wt, int Cout, int Cin) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * 9 * Cin) return;
  const float* W = params + (offsets ? offsets[blockIdx.y] : 0);
  const int co = idx % Cout, t = (idx / Cout) % 9, ci = idx / (9 * Cout);
  Wt[(size_t)blockIdx.y * Cout * 9 * Cin + idx] = W[(size_t)co * 9 * Cin + t * Cin + ci];
}
// Weights of `layers` layers ([128][K] row-major, layer l at base + (offsets ? offsets[l] : l*layer_stride)) -> per layer and
// K stage of 32 the two B tiles of the 3xTF32 GEMM, hi then lo, each [8 chunk planes][128 rows][16 B] with the plane pitch of
// the shared-memory layout: what gemm_tf32_kernel streams with bulk copies (GemmArgs::Bpack).
__global__ void __launch_bounds__(256) pack_b_kernel(const float* __restrict__ base, const long long* __restrict__ offsets, long long layer_stride,
                                                    int K, uint8_t* __restrict__ out) {
  pdl_enter();
  const int stages = K / kGemmKStage;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // one 16-byte chunk: (stage, plane, row)
  if (idx >= stages * 8 * 128) return;
  const int row = idx & 127, plane = (idx >> 7) & 7, st = idx >> 10, l = blockIdx.y;
  const float* W = base + (offsets ? offsets[l] : (long long)l * layer_stride);
  const float4 v = *reinterpret_cast<const float4*>(W + (size_t)row * K + st * kGemmKStage + plane * 4);
  const float4 h = make_float4(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                               __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
  uint8_t* o = out + ((size_t)l * stages + st) * (2 * kRegionA) + plane * kPlaneA + row * 16;
  *reinterpret_cast<float4*>(o) = h;
  *reinterpret_cast<float4*>(o + kRegionA) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
}

// planes float32 [B][5][cells] (board_to_input, neural_network.py:156-196) -> X0 [B*cells][8] (channels 5..7 zero)
__global__ void __launch_bounds__(256) planes_nhwc_kernel(const float* __restrict__ planes, float* __restrict__ X0, long long P, int cells) {
  pdl_enter();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long b = p / cells; const int cell = (int)(p % cells);
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = c < 5 ? planes[(size_t)(b * 5 + c) * cells + cell] : 0.f;
  float4* o = reinterpret_cast<float4*>(X0 + (size_t)p * 8);
  o[0] = make_float4(v[0], v[1], v[2], v[3]); o[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// out[c] (+)= sum_r X[r][c] (bias gradients).  One block per 32 columns x 256 rows; with more than one row chunk the
// chunk sums are added to the (zeroed) output with float atomics.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int ld, int R, int C, float* __restrict__ out, int atomic) {
  pdl_enter();
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * 256, r1 = min(R, r0 + 256);
  float s = 0.f;
  if (c < C) for (int r = r0 + ty; r < r1; r += 8) s += X[(size_t)r * ld + c];
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int j = 1; j < 8; ++j) s += part[j][tx];
    if (atomic) atomicAdd(out + c, s); else out[c] = s;
  }
}

// ------------------------------------------------------------------------------------------------ batch norm (training mode)
// nn.BatchNorm2d forward in train(): statistics over all P positions of the batch (biased variance for the
// normalisation, unbiased for running_var, momentum 0.1, eps 1e-5).  sums = float64 [2C]: forward sum / sum of squares;
// backward sum dZ*xhat (= d gamma) / sum dZ (= d beta), dZ = dOut * [Out > 0].
// 256 threads = (C/4 float4 column groups) x (1024/C row lanes); 32 rows per block
template <bool BWD>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ Y, int ldy, const float* __restrict__ dOut, int ldd,
                                                       const float* __restrict__ Out, int ldo, const float* __restrict__ mean_invstd,
                                                       int P, int C, double* __restrict__ sums) {
  pdl_enter();
  __shared__ float red[2][256][4];
  const int C4 = C >> 2, cg = threadIdx.x % C4, rl = threadIdx.x / C4, lanes = 256 / C4, c = cg * 4;
  const int r0 = blockIdx.x * 32, r1 = min(P, r0 + 32);
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  float mu[4] = {0.f, 0.f, 0.f, 0.f}, is[4] = {0.f, 0.f, 0.f, 0.f};
  if (BWD) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { mu[j] = mean_invstd[c + j]; is[j] = mean_invstd[C + c + j]; }
  }
  for (int r = r0 + rl; r < r1; r += lanes) {
    float y[4];
    *reinterpret_cast<float4*>(y) = *reinterpret_cast<const float4*>(Y + (size_t)r * ldy + c);
    if (!BWD) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[j] += y[j]; b[j] += y[j] * y[j]; }
    } else {
      float dz[4];
      *reinterpret_cast<float4*>(dz) = *reinterpret_cast<const float4*>(dOut + (size_t)r * ldd + c);
      if (Out) {
        float o[4];
        *reinterpret_cast<float4*>(o) = *reinterpret_cast<const float4*>(Out + (size_t)r * ldo + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) if (!(o[j] > 0.f)) dz[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[j] += dz[j] * ((y[j] - mu[j]) * is[j]); b[j] += dz[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[0][threadIdx.x][j] = a[j]; red[1][threadIdx.x][j] = b[j]; }
  __syncthreads();
  if (rl == 0) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[j] += red[0][l * C4 + cg][j]; b[j] += red[1][l * C4 + cg][j]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&sums[c + j], (double)a[j]); atomicAdd(&sums[C + c + j], (double)b[j]); }
  }
}
// out = [relu]( gamma * (Y - mean) * invstd + beta [+ residual] ), mean / invstd from the float64 sums; block 0 also writes
// mean_invstd float [2C] for the backward pass and updates running_mean / running_var in place (may be NULL).
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ Y, int ld, long long P, int C4, const double* __restrict__ sums,
                                                      float eps, float momentum, float* __restrict__ mean_invstd, float* __restrict__ running_mean,
                                                      float* __restrict__ running_var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ residual, int ldr, float* __restrict__ out, int ldo, int relu) {
  pdl_enter();
  __shared__ float s_mu[128], s_is[128];
  const int C = C4 * 4;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {         // the float64 arithmetic once per channel and block
    const double mean = sums[c] / P;
    double var = sums[C + c] / P - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mu = (float)mean, is = (float)(1.0 / sqrt(var + (double)eps));
    s_mu[c] = mu; s_is[c] = is;
    if (blockIdx.x == 0) {
      mean_invstd[c] = mu; mean_invstd[C + c] = is;
      if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
      if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(P > 1 ? var * P / (P - 1) : var);
    }
  }
  __syncthreads();
  const long long total = P * C4, i0 = (long long)blockIdx.x * 512 + threadIdx.x, i1 = i0 + 256;   // two float4 per thread, loads first
  const bool ok0 = i0 < total, ok1 = i1 < total;
  const long long p0 = i0 / C4, p1 = i1 / C4;
  const int c0 = (int)(i0 % C4) * 4, c1 = (int)(i1 % C4) * 4;
  float y0[4] = {0.f, 0.f, 0.f, 0.f}, y1[4] = {0.f, 0.f, 0.f, 0.f}, r0[4] = {0.f, 0.f, 0.f, 0.f}, r1[4] = {0.f, 0.f, 0.f, 0.f};
  if (ok0) *reinterpret_cast<float4*>(y0) = *reinterpret_cast<const float4*>(Y + (size_t)p0 * ld + c0);
  if (ok1) *reinterpret_cast<float4*>(y1) = *reinterpret_cast<const float4*>(Y + (size_t)p1 * ld + c1);
  if (residual) {
    if (ok0) *reinterpret_cast<float4*>(r0) = *reinterpret_cast<const float4*>(residual + (size_t)p0 * ldr + c0);
    if (ok1) *reinterpret_cast<float4*>(r1) = *reinterpret_cast<const float4*>(residual + (size_t)p1 * ldr + c1);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    y0[j] = (y0[j] - s_mu[c0 + j]) * s_is[c0 + j] * gamma[c0 + j] + beta[c0 + j] + r0[j];
    y1[j] = (y1[j] - s_mu[c1 + j]) * s_is[c1 + j] * gamma[c1 + j] + beta[c1 + j] + r1[j];
    if (relu) { y0[j] = fmaxf(y0[j], 0.f); y1[j] = fmaxf(y1[j], 0.f); }
  }
  if (ok0) *reinterpret_cast<float4*>(out + (size_t)p0 * ldo + c0) = *reinterpret_cast<float4*>(y0);
  if (ok1) *reinterpret_cast<float4*>(out + (size_t)p1 * ldo + c1) = *reinterpret_cast<float4*>(y1);
}
// backward, pass 2: dY = gamma*invstd*(dZ - dbeta/P - xhat*dgamma/P); optional dRes = dZ (the skip connection's share);
// block 0 also writes d gamma / d beta into the gradient buffer; dbias (optional, zero on entry) += column sums of dY --
// the gradient of the bias of the convolution in front of this batch norm (float atomics: its true value is zero, both
// this and the reference hold rounding noise there).
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dOut, int ldd, const float* __restrict__ Out, int ldo,
                                                          const float* __restrict__ Y, int ldy, const float* __restrict__ mean_invstd,
                                                          const float* __restrict__ gamma, const double* __restrict__ sums, long long P, int C4,
                                                          float* __restrict__ dY, int lddy, float* __restrict__ dRes, int lddr,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                                                          float* __restrict__ dYT, int ldt) {
  pdl_enter();
  __shared__ float red[256][4];
  __shared__ float tile[4608];                                 // the block's rows x (C+1): 32 x 129, 128 x 33 or 512 x 9
  const int C = C4 * 4;
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) { dgamma[c] = (float)sums[c]; dbeta[c] = (float)sums[C + c]; }
  const int cg = threadIdx.x % C4, c = cg * 4;                 // 256 % C4 == 0: a block is whole rows (kBnSlabs slabs of them)
  const float invP = 1.f / (float)P;
  float mu[4], is[4], ga[4], sg[4], sb[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mu[j] = mean_invstd[c + j]; is[j] = mean_invstd[C + c + j]; ga[j] = gamma[c + j];
    sg[j] = (float)sums[c + j] * invP; sb[j] = (float)sums[C + c + j] * invP;
  }
  float dzv[4][4], yv[4][4], ov[4][4];                           // all loads of the block's 4 row slabs first
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) {
    const long long idx = ((long long)blockIdx.x * 4 + sl) * 256 + threadIdx.x;
    const bool ok = idx < P * C4;
    const long long p = ok ? idx / C4 : 0;
    *reinterpret_cast<float4*>(dzv[sl]) = ok ? *reinterpret_cast<const float4*>(dOut + (size_t)p * ldd + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(yv[sl]) = ok ? *reinterpret_cast<const float4*>(Y + (size_t)p * ldy + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(ov[sl]) = (ok && Out) ? *reinterpret_cast<const float4*>(Out + (size_t)p * ldo + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) {
    const long long idx = ((long long)blockIdx.x * 4 + sl) * 256 + threadIdx.x;
    if (idx < P * C4) {
      const long long p = idx / C4;
      float r[4];
      float* dz = dzv[sl];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!(ov[sl][j] > 0.f)) dz[j] = 0.f;
        const float xhat = (yv[sl][j] - mu[j]) * is[j];
        r[j] = ga[j] * is[j] * (dz[j] - sb[j] - xhat * sg[j]);
        acc[j] += r[j];
      }
      *reinterpret_cast<float4*>(dY + (size_t)p * lddy + c) = *reinterpret_cast<float4*>(r);
      if (dRes) *reinterpret_cast<float4*>(dRes + (size_t)p * lddr + c) = *reinterpret_cast<float4*>(dz);
      if (dYT) {
        float* t = tile + (sl * (256 / C4) + threadIdx.x / C4) * (C + 1) + c;
        t[0] = r[0]; t[1] = r[1]; t[2] = r[2]; t[3] = r[3];
      }
    }
  }
  if (dYT) {                                                   // dYT[c][p] = dY[p][c]: 128 contiguous bytes per warp store
    __syncthreads();
    const int rb = 4 * (256 / C4);
    const long long p0 = (long long)blockIdx.x * rb;
    for (int cc = threadIdx.x >> 5; cc < C; cc += 8)
      for (int i = threadIdx.x & 31; i < rb; i += 32)
        if (p0 + i < P) dYT[(size_t)cc * ldt + p0 + i] = tile[i * (C + 1) + cc];
  }
  if (dbias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) red[threadIdx.x][j] = acc[j];
    __syncthreads();
    if (threadIdx.x < C4) {
      for (int l = 1; l < 256 / C4; ++l)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += red[l * C4 + cg][j];
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(dbias + c + j, acc[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ heads: losses and their gradients
// One warp per sample.  Policy: nn.CrossEntropyLoss with probability targets (trainer.py:61,131): loss_p = mean_b( -sum_a pi*log_softmax ),
// dlogits = (softmax * sum(pi) - pi) / B.  Value: v = tanh(relu(h) . w2 + b2) (neural_network.py:119-121), nn.MSELoss (trainer.py:60,132):
// loss_v = mean_b (v - z)^2; dpre = 2 (v - z)/B * (1 - v^2); dh = dpre * w2 * [h > 0] (h = value_fc1's output before its ReLU).
// losses[0] += policy loss, losses[1] += value loss (zeroed by the caller).
__global__ void __launch_bounds__(128) heads_loss_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ pi, int A,
                                                        const float* __restrict__ h, int ldh, int H, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, const float* __restrict__ z, int B,
                                                        float* __restrict__ dlogits, int lddl, float* __restrict__ dh, int lddh,
                                                        float* __restrict__ dpre_out, float* __restrict__ v_out, float* __restrict__ losses) {
  pdl_enter();
  const int lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* lg = logits + (size_t)b * ldl;
  const float* pb = pi + (size_t)b * A;
  float mx = -INFINITY;
  for (int a = lane; a < A; a += 32) mx = fmaxf(mx, lg[a]);
  for (int d = 16; d; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  float se = 0.f, sp = 0.f, spl = 0.f;
  for (int a = lane; a < A; a += 32) { se += expf(lg[a] - mx); sp += pb[a]; spl += pb[a] * (lg[a] - mx); }
  for (int d = 16; d; d >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, d); sp += __shfl_xor_sync(0xffffffffu, sp, d); spl += __shfl_xor_sync(0xffffffffu, spl, d);
  }
  const float lse = logf(se);
  const float invB = 1.f / (float)B;
  for (int a = lane; a < A; a += 32) dlogits[(size_t)b * lddl + a] = (expf(lg[a] - mx - lse) * sp - pb[a]) * invB;
  const float* hb = h + (size_t)b * ldh;
  float dot = 0.f;
  for (int j = lane; j < H; j += 32) dot += fmaxf(hb[j], 0.f) * w2[j];
  for (int d = 16; d; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
  const float v = tanhf(dot + b2[0]);
  const float err = v - z[b];
  const float dpre = 2.f * err * invB * (1.f - v * v);
  for (int j = lane; j < H; j += 32) dh[(size_t)b * lddh + j] = hb[j] > 0.f ? dpre * w2[j] : 0.f;
  if (lane == 0) {
    dpre_out[b] = dpre; v_out[b] = v;
    atomicAdd(&losses[0], (lse * sp - spl) * invB);
    atomicAdd(&losses[1], err * err * invB);
  }
}
// dw2[j] = sum_b dpre[b]*h[b][j]; db2 = sum_b dpre[b]
__global__ void value_fc2_grad_kernel(const float* __restrict__ dpre, const float* __restrict__ h, int ldh, int H, int B,
                                      float* __restrict__ dw2, float* __restrict__ db2) {
  pdl_enter();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < H) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dpre[b] * fmaxf(h[(size_t)b * ldh + j], 0.f);
    dw2[j] = s;
  }
  if (j == 0) { float s = 0.f; for (int b = 0; b < B; ++b) s += dpre[b]; db2[0] = s; }
}

// ------------------------------------------------------------------------------------------------ Adam (torch.optim.Adam, trainer.py:52-56)
// step_state: int step counter followed (at float index 2, 3) by lr/(1-beta1^t) and sqrt(1-beta2^t) of the current step
__global__ void adam_tick_kernel(int* step_state, float lr, float b1, float b2) {
  pdl_enter();
  const int t = ++step_state[0];
  float* f = reinterpret_cast<float*>(step_state);
  f[2] = (float)((double)lr / (1.0 - pow((double)b1, (double)t)));
  f[3] = (float)sqrt(1.0 - pow((double)b2, (double)t));
}
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                  long long n4, float b1, float b2, float eps, float wd, const int* __restrict__ step_state) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float step_size = reinterpret_cast<const float*>(step_state)[2], bc2s = reinterpret_cast<const float*>(step_state)[3];
  float pi[4], gi[4], mi[4], vi[4];
  *reinterpret_cast<float4*>(pi) = reinterpret_cast<const float4*>(p)[i]; *reinterpret_cast<float4*>(gi) = reinterpret_cast<const float4*>(g)[i];
  *reinterpret_cast<float4*>(mi) = reinterpret_cast<const float4*>(m)[i]; *reinterpret_cast<float4*>(vi) = reinterpret_cast<const float4*>(v)[i];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float gg = gi[j] + wd * pi[j];                      // L2 weight decay folded into the gradient
    mi[j] = mi[j] + (gg - mi[j]) * (1.f - b1);                // exp_avg.lerp_(grad, 1 - beta1)
    vi[j] = vi[j] * b2 + (1.f - b2) * gg * gg;
    pi[j] = pi[j] - step_size * (mi[j] / (sqrtf(vi[j]) / bc2s + eps));
  }
  reinterpret_cast<float4*>(p)[i] = *reinterpret_cast<float4*>(pi);
  reinterpret_cast<float4*>(m)[i] = *reinterpret_cast<float4*>(mi); reinterpret_cast<float4*>(v)[i] = *reinterpret_cast<float4*>(vi);
}

static long long* g_gemm_dbg = nullptr;

static int need_device() {
  int nd = 0;
  if (cudaGetDeviceCount(&nd) != cudaSuccess || nd == 0) { cudaGetLastError(); return set_error(YY_ERR_NO_DEVICE, "no CUDA device: the learner has no CPU fallback"); }
  return YY_OK;
}

}  // namespace yy

using namespace yy;

extern "C" {

int yy_lrn_gemm(const float* A, int lda, int a_mode, const float* B, int ldb, int b_mode, float* C, int ldc, int M, int N, int K,
                const float* bias, int relu, int accumulate, int tile_n, int split_k, float* ws, int64_t ws_floats, int precision,
                const yy_conv_geom* conv, const yy_gemm_stats* stats, const void* b_packed, void* stream) {
  int rc = need_device(); if (rc) return rc;
  if (M <= 0 || N <= 0 || K <= 0) return set_error(YY_ERR_INVALID, "gemm: empty problem");
  if ((lda | ldb | ldc | N | K) & 3) return set_error(YY_ERR_INVALID, "gemm: lda, ldb, ldc, N and K must be multiples of 4 floats");
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C | (uintptr_t)ws) & 15) return set_error(YY_ERR_INVALID, "gemm: operands must be 16-byte aligned");
  if (a_mode != YY_OP_K && a_mode != YY_OP_K_CONV) return set_error(YY_ERR_INVALID, "gemm: bad a_mode");
  if (b_mode != YY_OP_K && b_mode != YY_OP_K_CONVT) return set_error(YY_ERR_INVALID, "gemm: bad b_mode");
  if (a_mode == YY_OP_K_CONV || b_mode == YY_OP_K_CONVT) {
    if (!conv || conv->rows < 1 || conv->cols < 1 || conv->cin < 4 || (conv->cin & 3) || (a_mode == YY_OP_K_CONV && b_mode == YY_OP_K_CONVT))
      return set_error(YY_ERR_INVALID, "gemm: an implicit convolution operand (one per call) needs a geometry with cin a multiple of 4");
    if (a_mode == YY_OP_K_CONV && (K != 9 * conv->cin || M % (conv->rows * conv->cols))) return set_error(YY_ERR_INVALID, "gemm: conv A needs K = 9*cin and whole boards");
    if (b_mode == YY_OP_K_CONVT && (N != 9 * conv->cin || K % (conv->rows * conv->cols) || ((conv->rows * conv->cols) & 3)))
      return set_error(YY_ERR_INVALID, "gemm: conv B needs N = 9*cin and K = whole boards of a multiple of 4 cells");
  }
  if (precision != YY_GEMM_TF32 && precision != YY_GEMM_3XTF32) return set_error(YY_ERR_INVALID, "gemm: precision must be YY_GEMM_TF32 or YY_GEMM_3XTF32");
  if (tile_n < 16 || tile_n > 128 || tile_n % 16) return set_error(YY_ERR_INVALID, "gemm: tile_n in [16,128] step 16");
  if (split_k < 1) return set_error(YY_ERR_INVALID, "gemm: split_k >= 1");
  if (b_packed && (precision != YY_GEMM_3XTF32 || N != 128 || tile_n != 128 || (K % kGemmKStage) || b_mode != YY_OP_K || ((uintptr_t)b_packed & 15)))
    return set_error(YY_ERR_INVALID, "gemm: a packed B needs 3xTF32, N = tile_n = 128 and K a multiple of 32");
  double* bn_sums = stats ? stats->sums : nullptr;
  const bool bwd_stats = stats && stats->out != nullptr;
  if (bn_sums && (N > 128 || 128 % N)) return set_error(YY_ERR_INVALID, "gemm: fused batch-norm statistics need N dividing 128");
  if (bwd_stats && (!stats->y || !stats->mean_invstd || (stats->ldo & 3) || (stats->ldy & 3))) return set_error(YY_ERR_INVALID, "gemm: backward statistics need out, y and mean_invstd");
  int kps = (K + split_k - 1) / split_k;
  kps = (kps + kGemmKStage - 1) / kGemmKStage * kGemmKStage;
  const int zs = (K + kps - 1) / kps;
  if (b_mode == YY_OP_K_CONVT && kps > 8192) return set_error(YY_ERR_INVALID, "gemm: conv B supports at most 8192 positions per K slice (raise split_k)");
  const int edge_bytes = b_mode == YY_OP_K_CONVT ? (kps + 15) / 16 * 16 : 0;
  // K slices are summed on chip inside clusters of cz = 8, 4 or 2 CTAs (distributed shared memory) -- the largest size
  // whose clusters are all co-resident (one wave; a cluster needs its SMs in one GPC); what is left (groups = zs / cz > 1
  // slices) goes through the workspace and the ordered reducer.
  const int half = kRegionA + 8 * (tile_n * 16 + 16);
  const int smem_bytes = (precision == YY_GEMM_3XTF32 ? 3 * 2 * half : 4 * half) + edge_bytes;
  cudaStream_t st = (cudaStream_t)stream;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((M + 127) / 128), (unsigned)((N + tile_n - 1) / tile_n), (unsigned)zs);
  cfg.blockDim = dim3(288); cfg.stream = st; cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  {
    static int max_set[2] = {0, 0};
    const int v = precision == YY_GEMM_3XTF32 ? 1 : 0;
    if (smem_bytes > max_set[v]) {
      if (v) YY_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      else YY_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      max_set[v] = smem_bytes;
    }
  }
  int cz = 1;
  if (zs > 1 && tile_n % 4 == 0 && 256 % (tile_n / 4) == 0) {
    static int active[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};   // max co-resident clusters of 1 / 2 / 4 / 8 CTAs (largest tile), per precision
    const int v = precision == YY_GEMM_3XTF32 ? 1 : 0;
    const long long ctas = (long long)cfg.gridDim.x * cfg.gridDim.y * zs;
    for (int c = 8, i = 3; c >= 2; c >>= 1, --i) {
      if (zs % c) continue;
      if (!active[v][i]) {
        cudaLaunchConfig_t q = cfg;
        cudaLaunchAttribute qa[1] = {attr[0]};
        qa[0].val.clusterDim.z = (unsigned)c; q.attrs = qa; q.numAttrs = 1;
        q.dynamicSmemBytes = (size_t)smem_bytes;               // queried once per size class with the first caller's tile
        q.gridDim = dim3(1, 1, (unsigned)c);
        int n = 0;
        cudaError_t e = v ? cudaOccupancyMaxActiveClusters(&n, gemm_tf32_kernel<3, true>, &q) : cudaOccupancyMaxActiveClusters(&n, gemm_tf32_kernel<4, false>, &q);
        if (e != cudaSuccess) { cudaGetLastError(); n = -1; }
        active[v][i] = n > 0 ? n : -1;
      }
      if (active[v][i] > 0 && ctas / c <= active[v][i]) { cz = c; break; }
    }
  }
  attr[0].val.clusterDim.z = (unsigned)cz;
  const int groups = zs / cz;
  if (groups > 1 && (!ws || ws_floats < (int64_t)groups * M * N)) return set_error(YY_ERR_INVALID, "gemm: split-K beyond one cluster needs a workspace of (split/cluster)*M*N floats");
  const bool fuse_stats = bn_sums && cz > 1 && groups == 1 && N <= tile_n;
  GemmArgs g{A, B, C, bias, ws, (const uint8_t*)b_packed, lda, ldb, ldc, M, N, K, tile_n, kps, relu, accumulate, a_mode, b_mode,
             conv ? conv->rows : 1, conv ? conv->cols : 1, conv ? conv->cin : 4, conv ? conv->flip : 0, cz, fuse_stats ? bn_sums : nullptr,
             bwd_stats ? stats->out : nullptr, bwd_stats ? stats->y : nullptr, bwd_stats ? stats->mean_invstd : nullptr,
             bwd_stats ? stats->ldo : 0, bwd_stats ? stats->ldy : 0, g_gemm_dbg};
  if (precision == YY_GEMM_3XTF32) YY_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<3, true>, g));
  else YY_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<4, false>, g));
  YY_LAUNCH_CHECK();
  if (groups > 1) {
    if (256 % (N / 4) == 0) {
      const int rows_per_block = 4 * (256 / (N / 4));
      YY_CUDA_OK(launch_pdl(gemm_reduce_kernel, dim3((unsigned)((M + rows_per_block - 1) / rows_per_block)), dim3(256), 0, st, ws, groups, C, ldc, M, N / 4, bias, relu, accumulate,
                                                                                                   bwd_stats ? nullptr : bn_sums));
    } else {
      const long long total = (long long)M * (N / 4);
      YY_CUDA_OK(launch_pdl(gemm_reduce_flat_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, ws, groups, C, ldc, M, N / 4, bias, relu, accumulate));
    }
    YY_LAUNCH_CHECK();
  } else if (bn_sums && !fuse_stats && !bwd_stats) {
    YY_CUDA_OK(launch_pdl(bn_reduce_kernel<false>, dim3((M + 31) / 32), dim3(256), 0, st, C, ldc, nullptr, 0, nullptr, 0, nullptr, M, N, bn_sums));
    YY_LAUNCH_CHECK();
  }
  if (bwd_stats && !fuse_stats) {                             // not finished inside one cluster: a pass of its own over the finished C
    YY_CUDA_OK(launch_pdl(bn_reduce_kernel<true>, dim3((M + 31) / 32), dim3(256), 0, st, stats->y, stats->ldy, C, ldc, stats->out, stats->ldo, stats->mean_invstd, M, N, bn_sums));
    YY_LAUNCH_CHECK();
  }
  return YY_OK;
}

int yy_lrn_pack_b(const float* base, const long long* offsets_dev, int64_t layer_stride, int layers, int N, int K, void* out, void* stream) {
  int rc = need_device(); if (rc) return rc;
  if (N != 128 || K <= 0 || (K % kGemmKStage) || ((uintptr_t)out & 15)) return set_error(YY_ERR_INVALID, "pack_b: N = 128, K a multiple of 32");
  if (layers < 1) return YY_OK;
  const int chunks = (K / kGemmKStage) * 8 * 128;
  YY_CUDA_OK(launch_pdl(pack_b_kernel, dim3(dim3((unsigned)((chunks + 255) / 256), (unsigned)layers)), dim3(256), 0, (cudaStream_t)stream, base, offsets_dev, layer_stride, K, (uint8_t*)out));
  YY_LAUNCH_CHECK();
  return YY_OK;
}
int64_t yy_lrn_pack_b_bytes(int N, int K) { return N == 128 && K % kGemmKStage == 0 ? (int64_t)(K / kGemmKStage) * 2 * kRegionA : -1; }

int yy_lrn_gemm_debug_stamps(long long* dbg_dev) { g_gemm_dbg = dbg_dev; return YY_OK; }

int yy_lrn_transpose(const float* in, int ldi, float* out, int ldo, int R, int C, int batch, int64_t in_stride, int64_t out_stride, void* stream) {
  int rc = need_device(); if (rc) return rc;
  if (R <= 0 || C <= 0 || batch <= 0) return YY_OK;
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)((R + 31) / 32), (unsigned)batch);
  YY_CUDA_OK(launch_pdl(transpose_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, in, ldi, out, ldo, R, C, in_stride, out_stride));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_lrn_im2col_t(const float* X, int ldx, float* colT, int ldo, int64_t positions, int rows, int cols, int C, void* stream) {
  int rc = need_device(); if (rc) return rc;
  if (positions % (rows * cols)) return set_error(YY_ERR_INVALID, "im2col_t: positions must be whole boards");
  if (positions == 0) return YY_OK;
  dim3 grid((unsigned)((positions + 127) / 128), (unsigned)((C + 31) / 32), 9);
  YY_CUDA_OK(launch_pdl(im2col_t_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, X, ldx, colT, ldo, (int)positions, rows, cols, C));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_lrn_conv_weight_t(const float* params, const long long* offsets_dev, int layers, float* Wt, int Cout, int Cin, void* stream) {
  int rc = need_device(); if (rc) return rc;
  if (layers < 1) return YY_OK;
  const int total = Cout * 9 * Cin;
  YY_CUDA_OK(launch_pdl(conv_weight_t_kernel, dim3(dim3((unsigned)((total + 255) / 256), (unsigned)layers)), dim3(256), 0, (cudaStream_t)stream, params, offsets_dev, Wt, Cout, Cin));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_lrn_planes_nhwc(const float* planes, float* X0, int64_t boards, int cells, void* stream) {
  int rc = need_device(); if (rc) return rc;
  const long long P = boards * cells;
  if (P == 0) return YY_OK;
  YY_CUDA_OK(launch_pdl(planes_nhwc_kernel, dim3((unsigned)((P + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, planes, X0, P, cells));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_lrn_colsum(const float* X, int ld, int R, int C, float* out, void* stream) {
  int rc = need_device(); if (rc) return rc;
  if (R <= 0 || C <= 0) return YY_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = (R + 255) / 256;
  if (chunks > 1) YY_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
  YY_CUDA_OK(launch_pdl(colsum_kernel, dim3(dim3((unsigned)((C + 31) / 32), (unsigned)chunks)), dim3(256), 0, st, X, ld, R, C, out, chunks > 1));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

static int bn_shape_ok(int C) { return C >= 4 && C <= 128 && 128 % C == 0; }

int yy_lrn_bn_forward(const float* Y, int ld, int P, int C, const float* gamma, const float* beta, const float* residual, int ldr,
                      float* out, int ldo, int relu, float eps, float momentum, double* sums_ws, int have_sums, float* mean_invstd,
                      float* running_mean, float* running_var, void* stream) {
  int rc = need_device(); if (rc) return rc;
  if (!bn_shape_ok(C)) return set_error(YY_ERR_INVALID, "batch norm: C must divide 128");
  cudaStream_t st = (cudaStream_t)stream;
  if (!have_sums) {
    YY_CUDA_OK(launch_pdl(bn_reduce_kernel<false>, dim3((P + 31) / 32), dim3(256), 0, st, Y, ld, nullptr, 0, nullptr, 0, nullptr, P, C, sums_ws));
    YY_LAUNCH_CHECK();
  }
  const long long total = (long long)P * (C / 4);
  YY_CUDA_OK(launch_pdl(bn_apply_kernel, dim3((unsigned)((total + 511) / 512)), dim3(256), 0, st, Y, ld, P, C / 4, sums_ws, eps, momentum, mean_invstd, running_mean, running_var,
                                                                  gamma, beta, residual, ldr, out, ldo, relu));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_lrn_bn_backward(const float* dOut, int ldd, const float* Out, int ldo, const float* Y, int ldy, int P, int C,
                       const float* mean_invstd, const float* gamma, double* sums_ws, float* dY, int lddy, float* dRes, int lddr,
                       float* dgamma, float* dbeta, float* dbias, float* dYT, int ldt, int have_sums, void* stream) {
  int rc = need_device(); if (rc) return rc;
  if (!bn_shape_ok(C)) return set_error(YY_ERR_INVALID, "batch norm: C must divide 128");
  cudaStream_t st = (cudaStream_t)stream;
  if (!have_sums) {
    YY_CUDA_OK(launch_pdl(bn_reduce_kernel<true>, dim3((P + 31) / 32), dim3(256), 0, st, Y, ldy, dOut, ldd, Out, ldo, mean_invstd, P, C, sums_ws));
    YY_LAUNCH_CHECK();
  }
  const long long total = (long long)P * (C / 4);
  YY_CUDA_OK(launch_pdl(bn_bwd_apply_kernel, dim3((unsigned)((total + 1023) / 1024)), dim3(256), 0, st, dOut, ldd, Out, ldo, Y, ldy, mean_invstd, gamma, sums_ws, P, C / 4,
                                                                      dY, lddy, dRes, lddr, dgamma, dbeta, dbias, dYT, ldt));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_lrn_heads_loss(const float* logits, int ldl, const float* pi, int A, const float* h, int ldh, int H, const float* w2,
                      const float* b2, const float* z, int B, float* dlogits, int lddl, float* dh, int lddh, float* dpre,
                      float* v_out, float* dw2, float* db2, float* losses, void* stream) {
  int rc = need_device(); if (rc) return rc;
  if (B <= 0) return set_error(YY_ERR_INVALID, "heads: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  YY_CUDA_OK(cudaMemsetAsync(losses, 0, 2 * sizeof(float), st));
  YY_CUDA_OK(launch_pdl(heads_loss_kernel, dim3((B + 3) / 4), dim3(128), 0, st, logits, ldl, pi, A, h, ldh, H, w2, b2, z, B, dlogits, lddl, dh, lddh, dpre, v_out, losses));
  YY_LAUNCH_CHECK();
  YY_CUDA_OK(launch_pdl(value_fc2_grad_kernel, dim3((H + 127) / 128), dim3(128), 0, st, dpre, h, ldh, H, B, dw2, db2));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_lrn_adam(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                float weight_decay, int* step_state, void* stream) {
  int rc = need_device(); if (rc) return rc;
  if (n & 3) return set_error(YY_ERR_INVALID, "adam: the flat buffers must hold a multiple of 4 floats");
  cudaStream_t st = (cudaStream_t)stream;
  YY_CUDA_OK(launch_pdl(adam_tick_kernel, dim3(1), dim3(1), 0, st, step_state, lr, beta1, beta2));
  YY_LAUNCH_CHECK();
  YY_CUDA_OK(launch_pdl(adam_kernel, dim3((unsigned)((n / 4 + 255) / 256)), dim3(256), 0, st, params, grads, m, v, n / 4, beta1, beta2, eps, weight_decay, step_state));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

}  // extern "C"
