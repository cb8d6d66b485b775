// yy_nn.cu -- bf16 tcgen05 inference of the reference policy/value network on a batch of leaf boards.
//
// Network (src/yin_yang/ai/neural_network.py:39-123, eval mode, BatchNorm folded into the convolutions by
// the host packer):  board_to_input (:156-196) -> conv3x3(5->C)+ReLU -> blocks x [conv3x3+ReLU, conv3x3 +
// skip + ReLU] -> {policy: conv1x1(C->32)+ReLU -> FC(32A->A)} , {value: conv1x1(C->32)+ReLU -> FC(32A->256)
// +ReLU -> FC(256->1) -> tanh};  predict() applies softmax over all A logits (:152).
//
// Kernel 1  tower_kernel (this file's hot spot; >99 % of the FLOPs):
//   One persistent CTA per SM walks "groups" of boards.  A group is laid out as a flat list of up to 512
//   padded positions: each board contributes (n+1) x (m+1) positions -- its n x m cells plus one zero column
//   on the right and one zero row below -- so that the 3x3 tap (dy,dx) of EVERY position is the position
//   dy*(m+1)+dx further along the list and zero padding comes for free.  The whole residual tower runs with
//   the group's activations resident in shared memory ([C/8 chunks][616 rows][8 ch] bf16 = no-swizzle
//   K-major core matrices, so a tap shift is just a +16 B/row move of the UMMA descriptor start address);
//   only the weights stream in (16 KB stages, cp.async.bulk into a 4-deep mbarrier ring, shared by the 4
//   M=128 tiles of the group).  For 8-wide boards the MMA's 8-row groups are the board rows themselves
//   (descriptor SBO = 9*16 B skips the zero column): 7 boards per group, 87.5 % of the MMA rows are real cells.  Accumulators live in TMEM (4 tiles x 128 fp32 columns = all 512 columns).
//   The skip connection never touches shared memory: conv1's epilogue re-loads the block input into the
//   TMEM accumulator (tcgen05.st) before overwriting it in place, and conv2 accumulates on top of it.
//   Warp roles: warp 0 = weight producer, warp 1 = MMA issuer (one elected thread), warps 2-17 = epilogue
//   (one warp per tile x TMEM lane quarter, software-pipelined tcgen05.ld).
//   Roofline: tensor.  Algorithmic FLOPs per board: 2*A*(9*16*C + blocks*2*9*C*C + C*64)  (+ FC heads).
// Kernel 2/3  gemm_bf16_tn (yy_gemm.cu): policy FC and value FC1 on tensor cores over the whole batch.
// Kernel 4  heads_finish_kernel: softmax, value FC2 + tanh.
#include <cuda_bf16.h>

#include "yy_gemm.cuh"
#include "yy_nn.cuh"
#include "yy_ptx.cuh"
#include "yy_tower.cuh"

namespace yy {
using namespace ptx;

__global__ void __launch_bounds__(TW_THREADS, 1) tower_kernel(const TowerArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const TowerGeo& g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = 2 * g.blocks + 2;  // stem + 2*blocks tower convs + head conv
  int16_t* pos_p = reinterpret_cast<int16_t*>(smem + SM_POS);                       // M row -> padded position
  int16_t* pos_tab = reinterpret_cast<int16_t*>(smem + SM_POS + 128 * TW_MAXT * 2);    // M row -> board*256 + cell, or -1
  const uint32_t bar0 = smem_u32(smem + SM_BAR);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (TW_STAGES + s); };
  const uint32_t acc_full = bar0 + 8u * (2 * TW_STAGES), act_ready = bar0 + 8u * (2 * TW_STAGES + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM);

  // ---- one-time setup ----
  for (int i = tid; i < TW_CHUNKS * TW_ROWS * 4; i += TW_THREADS) reinterpret_cast<uint32_t*>(smem + SM_ACT)[i] = 0u;
  for (int i = tid; i < 128 * TW_MAXT; i += TW_THREADS) {
    int v = -1, p, b, y, x;
    if (g.row_aligned) {
      const int R = (i >> 7) * 16 + ((i & 127) >> 3);
      x = i & 7; p = R * g.pitch + x; b = R / g.rows_per_board; y = R % g.rows_per_board;
    } else {
      p = i; b = p / g.PB; const int rem = p % g.PB; y = rem / g.pitch; x = rem % g.pitch;
    }
    if (b < g.Gb && y < g.n && x < g.m) v = b * 256 + y * g.m + x;
    pos_p[i] = (int16_t)p;
    pos_tab[i] = (int16_t)v;
  }
  if (tid == 0) {
    for (int s = 0; s < TW_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(acc_full, 1);
    mbar_init(act_ready, TW_EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (a.dbg && tid == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); a.dbg[128 + 2 * blockIdx.x] = (long long)gt; }
  const uint32_t act_base = smem_u32(smem + SM_ACT);
  const uint32_t ring_base = smem_u32(smem + SM_RING);
  // this CTA's run of boards [run_lo, run_hi); groups of Gb boards, the last one possibly shorter -> fewer tiles
  const long long run_lo = (long long)blockIdx.x * a.boards_per_cta;
  const long long run_hi = (run_lo + a.boards_per_cta < a.count) ? run_lo + a.boards_per_cta : a.count;
  // (+pitch+1: the taps of the last real position read that far; rows beyond the group's tiles may be stale)
  auto tiles_for = [&](long long b0) {
    long long nb = run_hi - b0; if (nb > g.Gb) nb = g.Gb;
    int t = g.row_aligned ? (int)((nb * g.rows_per_board + 15) >> 4) : (int)((nb * g.PB + g.pitch + 1 + 127) >> 7);
    return t < g.T ? t : g.T;
  };

  if (warp == 0) {
    // =========================================================== weight producer (whole warp walks the loop with
    // warp-uniform state; one elected lane issues the bulk copies -- keeps everything on the uniform datapath)
    uint32_t it = 0;
    for (long long b0 = run_lo; b0 < run_hi; b0 += g.Gb) {
      for (int l = 0; l < L; ++l) {
        const LayerInfo li = layer_info(l, g.blocks);
        const uint8_t* src = a.conv_stream + li.stream_off;
        for (int j = 0; j < li.n_stages; ++j, ++it) {
          const uint32_t slot = it % TW_STAGES;
          if (it >= TW_STAGES) mbar_wait(empty_bar(slot), ((it / TW_STAGES) - 1) & 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(full_bar(slot), (uint32_t)li.stage_bytes);
            bulk_g2s(ring_base + slot * TW_STAGE_BYTES, src + (long long)j * li.stage_bytes, (uint32_t)li.stage_bytes, full_bar(slot));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer (whole warp runs the control flow, the
    // tcgen05.mma / commit instructions are issued by one elected lane; descriptors are base + constant deltas)
    uint32_t it = 0, act_phase = 0;
    const uint64_t kTileDelta = (uint64_t)g.tile_adv;            // next M=128 tile, in 16-byte rows
    constexpr uint64_t kK16DeltaA = (2u * TW_ROWS * 16u) >> 4;   // next K=16 slice: +2 channel chunks
    for (long long b0 = run_lo; b0 < run_hi; b0 += g.Gb) {
      const int T = tiles_for(b0);
      if (a.dbg && lane == 0 && (blockIdx.x == 0 || blockIdx.x == 100)) a.dbg[512 + (blockIdx.x ? 16 : 0) + (int)((b0 - run_lo) / g.Gb)] = clock64();
      for (int l = 0; l < L; ++l) {
        const LayerInfo li = layer_info(l, g.blocks);
        const bool preloaded = (l >= 1 && l <= 2 * g.blocks && (l & 1) == 0);  // conv2: accumulator holds the skip input
        const uint32_t idesc = idesc_bf16(128, li.N);
        const uint64_t k16_delta_b = (uint64_t)((2u * (uint32_t)li.N * 16u) >> 4);
        mbar_wait(act_ready, act_phase); act_phase ^= 1;
        tc_fence_after();
        if (a.dbg && blockIdx.x == 0 && b0 == run_lo && lane == 0) a.dbg[l * 4 + 0] = clock64();
        for (int j = 0; j < li.n_stages; ++j, ++it) {
          const uint32_t slot = it % TW_STAGES;
          int tapshift, chunk0;
          stage_info(l, j, g.blocks, g.pitch, tapshift, chunk0);
          mbar_wait(full_bar(slot), (it / TW_STAGES) & 1);
          tc_fence_after();
          const uint64_t ad0 = smem_desc(act_base + (uint32_t)((chunk0 * TW_ROWS + TW_PAD + tapshift) * 16), TW_ROWS * 16, (uint32_t)g.sbo_bytes);
          const uint64_t bd0 = smem_desc(ring_base + slot * TW_STAGE_BYTES, (uint32_t)li.N * 16, 128);
          const uint32_t acc0 = (preloaded || j > 0) ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < TW_MAXT; ++t) {
              if (t < T) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (k < li.nk16)
                    tc_mma_bf16(tmem_base + (uint32_t)(t * 128), ad0 + (uint64_t)t * kTileDelta + (uint64_t)k * kK16DeltaA,
                                bd0 + (uint64_t)k * k16_delta_b, idesc, k > 0 ? 1u : acc0);
                }
              }
            }
            tc_commit(empty_bar(slot));
          }
          __syncwarp();
        }
        if (elect_one()) tc_commit(acc_full);
        __syncwarp();
        if (a.dbg && blockIdx.x == 0 && b0 == run_lo && lane == 0) a.dbg[l * 4 + 1] = clock64();
      }
    }
  } else {
    // =========================================================== epilogue warps (2..17): warp -> (tile, lane quarter)
    const int ew = warp - 2;
    const int quarter = warp & 3;        // TMEM lanes a warp may touch: 32*(warp_id % 4) ..
    const int tile0 = ew >> 2;           // with 16 warps every tile of the group has its own 4 warps
    constexpr int kTileStride = TW_EPI_WARPS / 4;
    uint32_t acc_phase = 0;
    uint8_t* act = smem + SM_ACT;
    for (long long b0 = run_lo; b0 < run_hi; b0 += g.Gb) {
      const int T = tiles_for(b0);
      // ---- stem input planes (board_to_input, neural_network.py:156-196) for my rows ----
      for (int t = tile0; t < T; t += kTileStride) {
        const int mi = t * 128 + quarter * 32 + lane;
        const int p = pos_p[mi];
        const int info = pos_tab[mi];
        uint4 c0 = make_uint4(0, 0, 0, 0);
        const long long board = b0 + (info >= 0 ? (info >> 8) : 0);
        if (info >= 0 && board < run_hi) {
          const int cell = info & 255, y = cell / g.m, x = cell % g.m;
          const uint64_t* bb = a.black + board * g.W; const uint64_t* wb = a.white + board * g.W;
          auto bit = [&](const uint64_t* v, int c) { return (int)((v[c >> 6] >> (c & 63)) & 1ull); };
          const int isb = bit(bb, cell), isw = bit(wb, cell);
          int rc = 0, cc = 0;
          for (int xx = 0; xx < g.m; ++xx) { int c = y * g.m + xx; rc += bit(bb, c) | bit(wb, c); }
          for (int yy = 0; yy < g.n; ++yy) { int c = yy * g.m + x; cc += bit(bb, c) | bit(wb, c); }
          const float rf = (float)((double)rc / (double)g.m), cf = (float)((double)cc / (double)g.n);
          const float rf_hi = __bfloat162float(__float2bfloat16_rn(rf)), cf_hi = __bfloat162float(__float2bfloat16_rn(cf));
          // channels: 0 empty, 1 black, 2 white, 3 row fill, 4 col fill, 5/6 = bf16 residuals of 3/4 (same weights)
          c0.x = pack_bf16x2((isb | isw) ? 0.0f : 1.0f, isb ? 1.0f : 0.0f);
          c0.y = pack_bf16x2(isw ? 1.0f : 0.0f, rf_hi);
          c0.z = pack_bf16x2(cf_hi, rf - rf_hi);
          c0.w = pack_bf16x2(cf - cf_hi, 0.0f);
        }
        *reinterpret_cast<uint4*>(act + (size_t)(0 * TW_ROWS + TW_PAD + p) * 16) = c0;
        *reinterpret_cast<uint4*>(act + (size_t)(1 * TW_ROWS + TW_PAD + p) * 16) = make_uint4(0, 0, 0, 0);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(act_ready);

      for (int l = 0; l < L; ++l) {
        const bool is_head = (l == L - 1);
        const bool is_conv1 = (l >= 1 && l <= 2 * g.blocks && (l & 1) == 1);
        const float* bias = a.conv_bias + (size_t)l * TW_C;
        mbar_wait(acc_full, acc_phase); acc_phase ^= 1;
        tc_fence_after();
        if (a.dbg && blockIdx.x == 0 && b0 == run_lo && tid == 64) a.dbg[l * 4 + 2] = clock64();
        for (int t = tile0; t < T; t += kTileStride) {
          const int mi = t * 128 + quarter * 32 + lane;
          const int p = pos_p[mi];
          const int info = pos_tab[mi];
          const long long board = b0 + (info >= 0 ? (info >> 8) : 0);
          const bool real = info >= 0 && board < run_hi;
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * 128);
          uint8_t* rowp = act + (size_t)(TW_PAD + p) * 16;
          if (!is_head) {
            // one 16-column chunk: bias + ReLU (+ park the skip input in TMEM) -> two 16-byte channel chunks in place
            auto process = [&](const uint32_t (&r)[16], int cc) {
              uint4* d0 = reinterpret_cast<uint4*>(rowp + (size_t)(2 * cc) * TW_ROWS * 16);
              uint4* d1 = reinterpret_cast<uint4*>(rowp + (size_t)(2 * cc + 1) * TW_ROWS * 16);
              if (is_conv1) {  // skip connection: conv2 will accumulate on top of the block input
                const uint4 x0 = *d0, x1 = *d1;
                uint32_t xr[16];
                const uint32_t xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) { xr[2 * q] = xs[q] << 16; xr[2 * q + 1] = xs[q] & 0xffff0000u; }
                tc_st16(taddr + cc * 16, xr);
              }
              const float4* b4 = reinterpret_cast<const float4*>(bias + cc * 16);
              float v[16];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 bq = __ldg(b4 + q);
                v[4 * q + 0] = fmaxf(__uint_as_float(r[4 * q + 0]) + bq.x, 0.0f);
                v[4 * q + 1] = fmaxf(__uint_as_float(r[4 * q + 1]) + bq.y, 0.0f);
                v[4 * q + 2] = fmaxf(__uint_as_float(r[4 * q + 2]) + bq.z, 0.0f);
                v[4 * q + 3] = fmaxf(__uint_as_float(r[4 * q + 3]) + bq.w, 0.0f);
              }
              if (!real) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.0f;
              }
              *d0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
              *d1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
            };
            // software pipeline: the TMEM load of chunk c+1 is in flight while chunk c is processed
            uint32_t ra[16], rb[16];
            tc_ld16(taddr, ra);
#pragma unroll 1
            for (int cc = 0; cc < TW_C / 16; cc += 2) {
              tc_wait_ld();
              tc_ld16(taddr + (cc + 1) * 16, rb);
              process(ra, cc);
              tc_wait_ld();
              if (cc + 2 < TW_C / 16) tc_ld16(taddr + (cc + 2) * 16, ra);
              process(rb, cc + 1);
            }
            if (is_conv1) tc_wait_st();
          } else {
            __nv_bfloat16* dst = a.headfeat + (size_t)board * (TW_HEADC * g.A) + (info & 255);
#pragma unroll 1
            for (int cc = 0; cc < TW_HEADC / 16; ++cc) {
              uint32_t r[16];
              tc_ld16(taddr + cc * 16, r);
              tc_wait_ld();
              if (real) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  dst[(size_t)(cc * 16 + j) * g.A] = __float2bfloat16_rn(fmaxf(__uint_as_float(r[j]) + __ldg(bias + cc * 16 + j), 0.0f));
              }
            }
          }
        }
        tc_fence_before();
        if (a.dbg && blockIdx.x == 0 && b0 == run_lo && tid == 64) a.dbg[l * 4 + 3] = clock64();
        if (!is_head) { fence_proxy_async_smem(); mbar_arrive(act_ready); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (a.dbg && tid == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); a.dbg[129 + 2 * blockIdx.x] = (long long)gt; a.dbg[544 + (blockIdx.x & 255)] = clock64(); }
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// softmax over all A logits (neural_network.py:152) and value FC2 + tanh (:121).  One warp per board.
__global__ void __launch_bounds__(128) heads_finish_kernel(const float* __restrict__ logits_buf, int ld_logits,
                                                          const float* __restrict__ hidden, const float* __restrict__ w2,
                                                          const float* __restrict__ b2, long long count, int A,
                                                          float* __restrict__ policy, float* __restrict__ value,
                                                          float* __restrict__ logits_out) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= count) return;
  const float* lg = logits_buf + i * ld_logits;
  float mx = -INFINITY;
  for (int a = lane; a < A; a += 32) mx = fmaxf(mx, lg[a]);
  for (int off = 16; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float sum = 0.0f;
  for (int a = lane; a < A; a += 32) sum += expf(lg[a] - mx);
  for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  for (int a = lane; a < A; a += 32) {
    policy[i * A + a] = expf(lg[a] - mx) / sum;
    if (logits_out) logits_out[i * A + a] = lg[a];
  }
  float acc = 0.0f;
  for (int k = lane; k < 256; k += 32) acc += hidden[i * 256 + k] * w2[k];
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) value[i] = tanhf(acc + b2[0]);
}

// ------------------------------------------------------------------------------------------------ host side
int64_t nn_weight_bytes(int rows, int cols, int channels, int blocks) {
  if (!nn_geometry_ok(rows, cols) || channels < 1 || channels > TW_C || blocks < 0)
    return set_error(YY_ERR_INVALID, "network geometry unsupported: %dx%d, %d channels, %d blocks (need cols<=%d, channels<=%d)",
                     rows, cols, channels, blocks, TW_PAD - 2, TW_C);
  return weight_layout(rows, cols, blocks).total;
}

static void nn_scratch_layout(int rows, int cols, int max_boards, int64_t& headfeat, int64_t& logits, int64_t& hidden, int64_t& total) {
  const int A = rows * cols, a_pad = (A + 15) / 16 * 16;
  int64_t off = 0;
  headfeat = off; off = align256(off + (int64_t)max_boards * TW_HEADC * A * 2);
  logits = off; off = align256(off + (int64_t)max_boards * a_pad * 4);
  hidden = off; off = align256(off + (int64_t)max_boards * 256 * 4);
  total = off;
}

size_t nn_workspace_bytes(const yy_engine_config& cfg) {
  if (cfg.evaluator != YY_EVAL_NN) return 0;
  int64_t a, b, c, total;
  nn_scratch_layout(cfg.rows, cfg.cols, cfg.n_games * (cfg.leaves_per_step > 0 ? cfg.leaves_per_step : 1), a, b, c, total);
  return (size_t)total;
}

int nn_init(NNState& nn, const yy_engine_config& cfg, void* scratch) {
  nn = NNState{};
  nn.rows = cfg.rows; nn.cols = cfg.cols; nn.A = cfg.rows * cfg.cols; nn.W = words_for_cells(nn.A);
  nn.channels = cfg.nn_channels; nn.blocks = cfg.nn_blocks;
  nn.max_boards = cfg.n_games * (cfg.leaves_per_step > 0 ? cfg.leaves_per_step : 1); nn.device = cfg.device;
  if (cfg.evaluator != YY_EVAL_NN) return YY_OK;
  if (!nn_geometry_ok(cfg.rows, cfg.cols) || cfg.nn_channels > TW_C)
    return set_error(YY_ERR_INVALID, "network geometry unsupported: %dx%d, %d channels", cfg.rows, cfg.cols, cfg.nn_channels);
  nn.scratch = scratch; nn.scratch_bytes = (int64_t)nn_workspace_bytes(cfg);
  YY_CUDA_OK(cudaDeviceGetAttribute(&nn.num_sms, cudaDevAttrMultiProcessorCount, cfg.device));
  int major = 0;
  YY_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cfg.device));
  if (major != 10) return set_error(YY_ERR_NO_DEVICE, "the inference kernels are sm_100a-only (device reports sm_%d*)", major);
  YY_CUDA_OK(cudaFuncSetAttribute(tower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
  nn.attrs_set = true;
  return YY_OK;
}
static int nn_drain_events(NNState& nn) {
  for (int i = 0; i < nn.ev_used; ++i) {
    YY_CUDA_OK(cudaEventSynchronize(nn.ev[2 * i + 1]));
    float ms = 0.0f;
    YY_CUDA_OK(cudaEventElapsedTime(&ms, nn.ev[2 * i], nn.ev[2 * i + 1]));
    nn.tower_ms += ms;
  }
  nn.ev_used = 0;
  return YY_OK;
}
void nn_destroy(NNState& nn) {
  if (nn.ev) { for (int i = 0; i < 2 * nn.ev_cap; ++i) cudaEventDestroy(nn.ev[i]); delete[] nn.ev; nn.ev = nullptr; }
}
int nn_set_profiling(NNState& nn, int enable) {
  if (enable && !nn.ev) {
    nn.ev_cap = 1024;
    nn.ev = new cudaEvent_t[2 * nn.ev_cap];
    for (int i = 0; i < 2 * nn.ev_cap; ++i) YY_CUDA_OK(cudaEventCreate(&nn.ev[i]));
  }
  nn.profiling = enable != 0; nn.ev_used = 0; nn.tower_launches = 0; nn.tower_ms = 0.0; nn.tower_boards = 0;
  return YY_OK;
}
int nn_get_profile(NNState& nn, long long* launches, double* total_ms, long long* boards) {
  int rc = nn_drain_events(nn); if (rc) return rc;
  if (launches) *launches = nn.tower_launches;
  if (total_ms) *total_ms = nn.tower_ms;
  if (boards) *boards = nn.tower_boards;
  return YY_OK;
}

int nn_load_weights(NNState& nn, const void* weights, int64_t bytes) {
  int64_t need = nn_weight_bytes(nn.rows, nn.cols, nn.channels, nn.blocks);
  if (need < 0) return (int)need;
  if (!weights || bytes < need) return set_error(YY_ERR_INVALID, "weight image too small: %lld < %lld", (long long)bytes, (long long)need);
  if (((uintptr_t)weights & 255) != 0) return set_error(YY_ERR_INVALID, "weight image must be 256-byte aligned");
  nn.weights = weights; nn.weight_bytes = bytes;
  return YY_OK;
}

int nn_forward(NNState& nn, const uint64_t* black, const uint64_t* white, int64_t count, float* policy, float* value,
               float* logits, cudaStream_t s) {
  if (!nn.attrs_set) return set_error(YY_ERR_STATE, "engine was not created with the NN evaluator");
  if (!nn.weights) return set_error(YY_ERR_STATE, "no weights loaded (yy_engine_load_weights)");
  const WeightLayout wl = weight_layout(nn.rows, nn.cols, nn.blocks);
  const uint8_t* wimg = static_cast<const uint8_t*>(nn.weights);
  int64_t o_feat, o_logits, o_hidden, total;
  nn_scratch_layout(nn.rows, nn.cols, nn.max_boards, o_feat, o_logits, o_hidden, total);
  uint8_t* sc = static_cast<uint8_t*>(nn.scratch);
  __nv_bfloat16* headfeat = reinterpret_cast<__nv_bfloat16*>(sc + o_feat);
  float* logits_buf = reinterpret_cast<float*>(sc + o_logits);
  float* hidden = reinterpret_cast<float*>(sc + o_hidden);
  const int A = nn.A;
  for (int64_t done = 0; done < count; done += nn.max_boards) {
    const int64_t n = (count - done) < nn.max_boards ? (count - done) : nn.max_boards;
    TowerArgs ta;
    ta.g = make_tower_geo(nn.rows, nn.cols, nn.blocks);
    ta.conv_stream = wimg + wl.conv_stream;
    ta.conv_bias = reinterpret_cast<const float*>(wimg + wl.conv_bias);
    ta.black = black + done * nn.W; ta.white = white + done * nn.W;
    ta.count = n;
    ta.headfeat = headfeat;
    ta.dbg = nn.dbg;
    // deal boards to CTAs in equal contiguous runs (wave balance: every SM gets the same number of boards; a run's
    // last group is short and uses fewer tiles), but never fewer than one full group per CTA
    long long per = (n + nn.num_sms - 1) / nn.num_sms;
    if (per < ta.g.Gb) per = ta.g.Gb;
    ta.boards_per_cta = (int)per;
    const int grid = (int)((n + per - 1) / per);
    if (nn.profiling) {
      if (nn.ev_used == nn.ev_cap) { int rc = nn_drain_events(nn); if (rc) return rc; }
      YY_CUDA_OK(cudaEventRecord(nn.ev[2 * nn.ev_used], s));
    }
    tower_kernel<<<grid, TW_THREADS, SM_TOTAL, s>>>(ta);
    YY_LAUNCH_CHECK();
    if (nn.profiling) {
      YY_CUDA_OK(cudaEventRecord(nn.ev[2 * nn.ev_used + 1], s));
      ++nn.ev_used; ++nn.tower_launches; nn.tower_boards += n;
    }
    GemmArgs gp{headfeat, TW_HEADC * A, reinterpret_cast<const __nv_bfloat16*>(wimg + wl.fc_policy_w), 32 * A, logits_buf, wl.a_pad,
                reinterpret_cast<const float*>(wimg + wl.fc_policy_b), (int)n, wl.a_pad, 32 * A, 0};
    GemmArgs gv{headfeat + 32 * A, TW_HEADC * A, reinterpret_cast<const __nv_bfloat16*>(wimg + wl.fc_value1_w), 32 * A, hidden, 256,
                reinterpret_cast<const float*>(wimg + wl.fc_value1_b), (int)n, 256, 32 * A, 1};
    int rc;
    if (wl.a_pad % 64 == 0) rc = gemm_bf16_tn_pair(gp, gv, s);          // policy FC + value FC1 in one launch
    else { rc = gemm_bf16_tn(gp, s); if (rc) return rc; rc = gemm_bf16_tn(gv, s); }
    if (rc) return rc;
    const unsigned grid2 = (unsigned)((n * 32 + 127) / 128);
    heads_finish_kernel<<<grid2, 128, 0, s>>>(logits_buf, wl.a_pad, hidden, reinterpret_cast<const float*>(wimg + wl.fc_value2_w),
                                              reinterpret_cast<const float*>(wimg + wl.fc_value2_b), n, A,
                                              policy + done * A, value + done, logits ? logits + done * A : nullptr);
    YY_LAUNCH_CHECK();
  }
  return YY_OK;
}

}  // namespace yy

// Section offsets of the packed weight image, for the host-side packer:
// out[0..8] = conv_stream, conv_bias, fc_policy_w, fc_policy_b, fc_value1_w, fc_value1_b, fc_value2_w, fc_value2_b, total;
// out[9] = padded policy rows (A rounded up to 16); out[10], out[11] = offset / bytes of the stage-ordered FC stream;
// out[12] = conv stream with N-halved stages (CTA-pair kernel).
extern "C" int yy_nn_weight_layout(int rows, int cols, int channels, int blocks, int64_t* out) {
  using namespace yy;
  int64_t t = nn_weight_bytes(rows, cols, channels, blocks);
  if (t < 0) return (int)t;
  WeightLayout w = weight_layout(rows, cols, blocks);
  out[0] = w.conv_stream; out[1] = w.conv_bias; out[2] = w.fc_policy_w; out[3] = w.fc_policy_b; out[4] = w.fc_value1_w;
  out[5] = w.fc_value1_b; out[6] = w.fc_value2_w; out[7] = w.fc_value2_b; out[8] = w.total; out[9] = w.a_pad;
  out[10] = w.fc_stream; out[11] = w.fc_stream_bytes; out[12] = w.conv_stream_pair;
  return YY_OK;
}
