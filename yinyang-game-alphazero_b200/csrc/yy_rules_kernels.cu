// yy_rules_kernels.cu -- batched, stateless rules entry points of the C ABI (one thread per board; the fused env step:
// two lanes per board).  Replaces YinYangGame.getValidMoves / getNextState / getGameEnded
// (src/yin_yang/yin_yang_game.py:39-110) over YinYangLogic (src/yin_yang/yin_yang_logic.py:24-134).
// Square one-word boards under the Python rules (6x6, 8x8) run the line-fill rules of yy_rules_sq.cuh, every other
// geometry and the row/column rule the generic bitboard rules of yy_rules.cuh.
//
// Roofline: HBM.  Algorithmic bytes per env step = 6*ceil(A/8)+4 (52 B at 8x8).  8x8, 65,536 boards: one
// launch is 5.4-5.9 us, of which 2.0 us are the floor of a one-block launch (launch + completion + a cold DRAM round trip)
// and the rest integer issue (1.79 M warp instructions); with 1 M boards the kernel is integer-issue bound at 22-26 G steps/s
// (profiles/r02_env_step_sq_ncu_summary.txt).
#include "yy_common.cuh"
#include "yy_rules_sq.cuh"

namespace yy {

static thread_local char t_err[512];
char* error_buffer() { return t_err; }
int set_error(int status, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(t_err, sizeof(t_err), fmt, ap); va_end(ap);
  return status;
}
std::atomic<long long> g_launches{0};

constexpr int kRulesBlock = 128;

template <int NW>
__global__ void __launch_bounds__(kRulesBlock)
legal_mask_kernel(Geo<NW> g, int W, const uint64_t* __restrict__ black, const uint64_t* __restrict__ white,
                  const int8_t* __restrict__ players, uint64_t* __restrict__ out_mask, long long count) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  BB<NW> b = load_bb<NW>(black, i, W) & g.full, w = load_bb<NW>(white, i, W) & g.full;
  store_bb<NW>(out_mask, i, W, legal_for(g, b, w, players[i] == 1 ? 1 : -1));
}

template <int NW>
__global__ void __launch_bounds__(kRulesBlock)
step_kernel(Geo<NW> g, int W, uint64_t* __restrict__ black, uint64_t* __restrict__ white, int8_t* __restrict__ players,
            const int32_t* __restrict__ actions, long long count) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  BB<NW> b = load_bb<NW>(black, i, W) & g.full, w = load_bb<NW>(white, i, W) & g.full;
  int p = players[i] == 1 ? 1 : -1;
  if (apply_action(g, b, w, p, actions[i])) {
    if (p == 1) store_bb<NW>(black, i, W, b); else store_bb<NW>(white, i, W, w);
  }
  players[i] = (int8_t)(-players[i]);  // yin_yang_game.py:58 returns -player whatever it was
}

template <int NW>
__global__ void __launch_bounds__(kRulesBlock)
ended_kernel(Geo<NW> g, int W, const uint64_t* __restrict__ black, const uint64_t* __restrict__ white,
             const int8_t* __restrict__ players, int8_t* __restrict__ out, long long count) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  BB<NW> b = load_bb<NW>(black, i, W) & g.full, w = load_bb<NW>(white, i, W) & g.full;
  out[i] = (int8_t)ended_code(g, b, w, players[i] == 1 ? 1 : -1);
}

// One-word square boards without the row/column rule: the line-fill rules of yy_rules_sq.cuh, one thread per board.
template <int SIDE> __device__ __forceinline__ int sq_ended_code(uint64_t mover, uint64_t other, uint64_t mover_mask) {
  if (mover_mask || Sq<SIDE>::legal(other, mover)) return 0;
  const int mc = popc64(mover), oc = popc64(other);
  return mc > oc ? 1 : (oc > mc ? -1 : 2);
}
template <int SIDE>
__global__ void __launch_bounds__(kRulesBlock)
legal_mask_sq_kernel(const uint64_t* __restrict__ black, const uint64_t* __restrict__ white, const int8_t* __restrict__ players,
                     uint64_t* __restrict__ out_mask, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t b = black[i] & Sq<SIDE>::FULL, w = white[i] & Sq<SIDE>::FULL;
  out_mask[i] = players[i] == 1 ? Sq<SIDE>::legal(b, w) : Sq<SIDE>::legal(w, b);
}
template <int SIDE>
__global__ void __launch_bounds__(kRulesBlock)
step_sq_kernel(uint64_t* __restrict__ black, uint64_t* __restrict__ white, int8_t* __restrict__ players,
               const int32_t* __restrict__ actions, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t b = black[i] & Sq<SIDE>::FULL, w = white[i] & Sq<SIDE>::FULL;
  const int praw = players[i], a = actions[i];
  if ((unsigned)a < (unsigned)Sq<SIDE>::CELLS) {
    const uint64_t bit = 1ull << a;
    const uint64_t lm = praw == 1 ? Sq<SIDE>::legal(b, w) : Sq<SIDE>::legal(w, b);
    if (lm & bit) (praw == 1 ? black : white)[i] = (praw == 1 ? b : w) | bit;
  }
  players[i] = (int8_t)(-praw);  // yin_yang_game.py:58 returns -player whatever it was
}
template <int SIDE>
__global__ void __launch_bounds__(kRulesBlock)
ended_sq_kernel(const uint64_t* __restrict__ black, const uint64_t* __restrict__ white, const int8_t* __restrict__ players,
                int8_t* __restrict__ out, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t b = black[i] & Sq<SIDE>::FULL, w = white[i] & Sq<SIDE>::FULL;
  const bool mover_black = players[i] == 1;
  const uint64_t p = mover_black ? b : w, o = mover_black ? w : b;
  out[i] = (int8_t)sq_ended_code<SIDE>(p, o, Sq<SIDE>::legal(p, o));
}

// Fused getValidMoves + getNextState + getGameEnded.
// Two lanes per board: lane 2k analyses the mover's colour, lane 2k+1 the other colour, so the two flood fills of a
// step (the only long dependent chains) run side by side and the chain per lane halves.  What a step needs is then
// exactly one component analysis per colour: the successor's masks follow without further fills, because the
// opponent's stones do not change and a legal placement leaves the mover's stones one component (its dilation is
// the new "touches every component" set).  The fills make the work per board data dependent (it grows with the
// number of stones) and a batch usually mixes game stages: in input order only 9 of 32 lanes were active per issued
// instruction (ncu, BASELINE configs[1]).  Each block therefore first counting-sorts its boards by stone count in
// shared memory, so the lanes of a warp get boards of similar cost; results go back to the boards' own slots.
constexpr int kEnvMaxBins = 64;

// SIDE = n for the common square boards (6, 8, 16): rows, cols and cells become compile-time constants, so the
// row shifts of the dilations are immediate funnel shifts (a 64-bit shift by a run-time count costs three times as many
// instructions) and the per-word loops have known trip counts; SIDE = 0 reads the geometry from g.
template <int NW, int BLOCK, bool SORT, int SIDE>
__global__ void __launch_bounds__(BLOCK)
env_step_kernel(Geo<NW> g, int W, uint64_t* __restrict__ black, uint64_t* __restrict__ white,
                int8_t* __restrict__ players, const int32_t* __restrict__ actions, uint64_t* __restrict__ out_mask,
                int8_t* __restrict__ out_result, long long count) {
  constexpr int BOARDS = BLOCK / 2;
  constexpr int kEnvBins = BOARDS > kEnvMaxBins ? kEnvMaxBins : 32;   // + one bin for the boards past the end
  static_assert(BLOCK > kEnvBins && kEnvBins % 32 == 0, "one thread per histogram bin");
  if (SIDE) {
    g.rows = SIDE; g.cols = SIDE; g.cells = SIDE * SIDE;
    if (SIDE * SIDE == 64 * NW) for (int k = 0; k < NW; ++k) g.full.w[k] = ~0ull;
  }
  if (NW <= 2) W = NW;                          // words_for_cells == NW unless the board has 129..192 cells
  __shared__ int s_hist[kEnvBins + 1];
  __shared__ int s_off[kEnvBins + 1];
  __shared__ uint16_t s_order[BOARDS];        // sorted position -> board of this block
  // everything a board needs is fetched in ONE round trip to DRAM, before the sort, and staged in shared memory: the
  // lanes that analyse the board after the sort read it from there (a second dependent miss for players / actions
  // would sit on the critical path of a 6 us kernel)
  __shared__ uint64_t s_black[SORT ? BOARDS * NW : 1], s_white[SORT ? BOARDS * NW : 1];
  __shared__ int32_t s_action[SORT ? BOARDS : 1];
  __shared__ int8_t s_player[SORT ? BOARDS : 1];
  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * BOARDS;
  if (SORT) {
    BB<NW> mb = bb_zero<NW>(), mw = bb_zero<NW>();
    int mp = 1, ma = -1;
    const bool mlive = tid < BOARDS && base + tid < count;
    if (mlive) {
      const long long mine = base + tid;
      mb = load_bb<NW>(black, mine, W); mw = load_bb<NW>(white, mine, W);
      mp = players[mine]; ma = actions[mine];
    }
    if (tid <= kEnvBins) s_hist[tid] = 0;
    __syncthreads();
    int key = kEnvBins, rank = 0;               // boards past the end sort last
    if (tid < BOARDS) {
      if (mlive) key = popcount(mb | mw) * kEnvBins / (g.cells + 1);
      rank = atomicAdd(&s_hist[key], 1);
  #pragma unroll
      for (int k = 0; k < NW; ++k) { s_black[tid * NW + k] = mb.w[k]; s_white[tid * NW + k] = mw.w[k]; }
      s_player[tid] = (int8_t)mp; s_action[tid] = ma;
    }
    __syncthreads();
    if (tid < 32) {                             // exclusive scan of the kEnvBins + 1 bins by one warp, 32 bins at a time
      int run = 0;
  #pragma unroll
      for (int seg = 0; seg <= kEnvBins / 32; ++seg) {
        const int idx = seg * 32 + tid;
        const int v = idx <= kEnvBins ? s_hist[idx] : 0;
        int inc = v;
  #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, d);
          if (tid >= d) inc += t;
        }
        if (idx <= kEnvBins) s_off[idx] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    __syncthreads();
    if (tid < BOARDS) s_order[s_off[key] + rank] = (uint16_t)tid;
    __syncthreads();
  }

  const int side = tid & 1;                   // 0: the mover's colour, 1: the other colour
  const int loc = SORT ? (int)s_order[tid >> 1] : (tid >> 1);
  const long long i = base + loc;
  const bool live = i < count;                // both lanes of a pair agree; nobody leaves before the shuffles
  BB<NW> b = bb_zero<NW>(), w = bb_zero<NW>();
  int praw = 1, a = -1;
  if (live) {
    if (SORT) {
  #pragma unroll
      for (int k = 0; k < NW; ++k) { b.w[k] = s_black[loc * NW + k]; w.w[k] = s_white[loc * NW + k]; }
      b = b & g.full; w = w & g.full;
      praw = s_player[loc]; a = s_action[loc];
    } else {
      b = load_bb<NW>(black, i, W) & g.full; w = load_bb<NW>(white, i, W) & g.full;
      praw = players[i]; a = actions[i];
    }
  }
  const int p = praw == 1 ? 1 : -1;
  const bool mine_black = (p == 1) == (side == 0);
  BB<NW> x = mine_black ? b : w, y = mine_black ? w : b;   // x: the colour this lane analyses
  const bool rowcol = (g.rule_flags & YY_RULE_ROWCOL_BIT) != 0;

  BB<NW> cand = legal_candidates(g, x, y);
  BB<NW> acc = g.full;                        // cells touching every component of x
  if (any(cand) || rowcol) acc = touches_all_components(g, x);
  BB<NW> lm = cand & acc;
  int placed = side == 0 && a >= 0 && a < g.cells && test(lm, a);
  placed = __shfl_sync(0xffffffffu, placed, (tid & 31) & ~1);
  if (live && side == 0) store_bb<NW>(out_mask, i, W, lm);
  if (placed) {                               // the pair takes this branch together
    if (side == 0) {
      setbit(x, a); acc = dilate4(g, x);
      store_bb<NW>(p == 1 ? black : white, i, W, x);
    } else {
      setbit(y, a);
    }
    lm = legal_candidates(g, x, y) & acc;     // masks of the successor
  }
  const int mine_any = any(lm);
  const int peer_any = __shfl_xor_sync(0xffffffffu, mine_any, 1);
  if (live && side == 0) {
    int code = 0;                             // getGameEnded(successor, -p)
    if (!mine_any && !peer_any) {
      int bc = popcount(p == 1 ? x : y), wc = popcount(p == 1 ? y : x);
      code = bc > wc ? (p == 1 ? -1 : 1) : (wc > bc ? (p == 1 ? 1 : -1) : 2);
    }
    out_result[i] = (int8_t)code;
    players[i] = (int8_t)(-praw);             // yin_yang_game.py:58 returns -player whatever it was
  }
}

// The same step for square one-word boards (6x6, 8x8) without the row/column rule: the rules of yy_rules_sq.cuh (line
// fills, pair-based 2x2 test), two lanes per board, boards in input order.  Nothing is staged or sorted: a line fill
// converges in 1..5 rounds whatever the number of stones, so the lanes of a warp finish close together without it.
// What each lane does after its colour's analysis is as short as the rules allow: a legal placement cannot create a
// 2x2 block, leaves the mover's stones ONE component (its dilation is the successor's "touches every component" set)
// and only removes a cell from the other colour's mask.
// The kernel is launched with programmatic stream serialisation: it lets the next launch in the stream start its blocks
// at once (griddepcontrol.launch_dependents) and itself waits for the previous kernel's memory (griddepcontrol.wait)
// before the first load, so back-to-back steps overlap launch latency with the previous step's execution.
template <int SIDE, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
env_step_sq_kernel(uint64_t* __restrict__ black, uint64_t* __restrict__ white, int8_t* __restrict__ players,
                   const int32_t* __restrict__ actions, uint64_t* __restrict__ out_mask, int8_t* __restrict__ out_result,
                   long long count) {
  using Q = Sq<SIDE>;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const unsigned tid = threadIdx.x;
  const unsigned side = tid & 1;              // 0: the mover's colour, 1: the other colour
  const unsigned i = (blockIdx.x * (unsigned)BLOCK + tid) >> 1;   // the host keeps 2 * count below 2^32 per launch
  const bool live = i < (unsigned)count;      // both lanes of a pair agree; nobody leaves before the shuffles
  asm volatile("griddepcontrol.wait;" ::: "memory");
  uint64_t b = 0, w = 0;
  int praw = 1, a = -1;
  if (live) { b = black[i] & Q::FULL; w = white[i] & Q::FULL; praw = players[i]; a = actions[i]; }
  const bool mover_black = praw == 1;
  const bool mine_black = mover_black == (side == 0);
  uint64_t x = mine_black ? b : w;            // the colour this lane analyses
  const uint64_t y = mine_black ? w : b;
  bool has; uint64_t completes;
  Q::blocks(x, has, completes);
  const bool any_block = __shfl_xor_sync(0xffffffffu, (int)has, 1) | (int)has;
  uint64_t empty = Q::FULL & ~(x | y);
  uint64_t lm = any_block ? 0ull : (empty & ~completes);
  if (lm) lm &= Q::touches_all(x, lm);
  const bool valid_action = (unsigned)a < (unsigned)Q::CELLS;
  const uint64_t abit = valid_action ? (1ull << (a & 63)) : 0ull;
  int placed = side == 0 && (lm & abit) != 0;
  placed = __shfl_sync(0xffffffffu, placed, (tid & 31) & ~1);
  if (live && side == 0) out_mask[i] = lm;
  if (placed) {                               // the pair takes this branch together
    if (side == 0) {
      x |= abit;
      (mover_black ? black : white)[i] = x;
      Q::blocks(x, has, completes);
      lm = empty & ~abit & ~completes & Q::dilate4(x);
    } else {
      lm &= ~abit;
    }
  }
  const int mine_any = lm != 0;
  const int peer_any = __shfl_xor_sync(0xffffffffu, mine_any, 1);
  if (live && side == 0) {
    int code = 0;                             // getGameEnded(successor, -p)
    if (!mine_any && !peer_any) {
      const int mc = popc64(x), oc = popc64(y);   // mover's stones (after the move) / the other colour's
      code = mc > oc ? -1 : (oc > mc ? 1 : 2);    // seen from the player to move next, i.e. the other colour
    }
    out_result[i] = (int8_t)code;
    players[i] = (int8_t)(-praw);             // yin_yang_game.py:58 returns -player whatever it was
  }
}

template <int NW>
__global__ void __launch_bounds__(kRulesBlock)
random_playout_kernel(Geo<NW> g, int W, uint64_t seed, const int32_t* __restrict__ plies, uint64_t* __restrict__ black,
                      uint64_t* __restrict__ white, int8_t* __restrict__ players, long long count) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  BB<NW> b = bb_zero<NW>(), w = bb_zero<NW>();
  int p = 1;
  Philox rng(seed, (uint64_t)i, 0x706c6179ull);
  int n = plies[i];
  for (int k = 0; k < n; ++k) {
    BB<NW> lm = legal_for(g, b, w, p);
    int c = popcount(lm);
    if (c) {
      int a = kth_bit(lm, (int)rng.below((uint32_t)c));
      if (p == 1) setbit(b, a); else setbit(w, a);
    }
    p = -p;  // a side without a legal move passes
  }
  store_bb<NW>(black, i, W, b); store_bb<NW>(white, i, W, w);
  players[i] = (int8_t)p;
}

// int8 boards (0 / +1 / -1, yin_yang_logic.py:14-22) <-> bitboards, on the device: the host-buffer entry points copy the
// reference's own board arrays and leave the bit packing to the GPU.  One thread per (board, word).
__global__ void __launch_bounds__(256) pack_boards_kernel(const int8_t* __restrict__ boards, int cells, int W, uint64_t* __restrict__ black,
                                                         uint64_t* __restrict__ white, long long count) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count * W) return;
  const long long i = idx / W; const int w = (int)(idx % W);
  const int8_t* b = boards + i * cells + w * 64;
  const int n = min(64, cells - w * 64);
  uint64_t bk = 0, wh = 0;
  for (int k = 0; k < n; ++k) { const int v = b[k]; bk |= (uint64_t)(v == 1) << k; wh |= (uint64_t)(v == -1) << k; }
  black[idx] = bk; white[idx] = wh;
}
// out_boards (optional) int8 [count][cells] = black - white; out_mask (optional) uint8 [count][cells] = bits of `mask`.  One thread per cell.
__global__ void __launch_bounds__(256) unpack_boards_kernel(const uint64_t* __restrict__ black, const uint64_t* __restrict__ white,
                                                           const uint64_t* __restrict__ mask, int cells, int W, int8_t* __restrict__ out_boards,
                                                           uint8_t* __restrict__ out_mask, long long count) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count * cells) return;
  const long long i = idx / cells; const int a = (int)(idx % cells);
  const long long wi = i * W + (a >> 6); const int bit = a & 63;
  if (out_boards) out_boards[idx] = (int8_t)((int)((black[wi] >> bit) & 1ull) - (int)((white[wi] >> bit) & 1ull));
  if (out_mask) out_mask[idx] = (uint8_t)((mask[wi] >> bit) & 1ull);
}

// side of a square one-word board that the line-fill kernels cover (6 or 8, Python rules only), else 0
static int sq_side(int rows, int cols, uint32_t rule_flags) {
  return rows == cols && (rows == 6 || rows == 8) && !(rule_flags & YY_RULE_ROWCOL_BIT) ? rows : 0;
}
static int check_rules_args(int rows, int cols, long long count) {
  if (!board_supported(rows, cols)) return set_error(YY_ERR_INVALID, "unsupported board %dx%d (need <=32 per side, <=256 cells)", rows, cols);
  if (count < 0) return set_error(YY_ERR_INVALID, "negative count");
  if (yy_device_count() == 0) return set_error(YY_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
  return YY_OK;
}

}  // namespace yy

using namespace yy;

extern "C" {

int yy_abi_version(void) { return YY_ABI_VERSION; }
const char* yy_last_error(void) { return yy::error_buffer(); }
int yy_device_count(void) {
  static std::atomic<int> seen{0};              // a positive answer is final: every entry point asks, once is enough
  int n = seen.load(std::memory_order_relaxed);
  if (n > 0) return n;
  n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (n > 0) seen.store(n, std::memory_order_relaxed);
  return n;
}
int64_t yy_launch_count(void) { return (int64_t)yy::g_launches.load(); }

int yy_legal_mask(int rows, int cols, uint32_t rule_flags, const uint64_t* black, const uint64_t* white,
                  const int8_t* players, uint64_t* out_mask, int64_t count, void* stream) {
  int rc = check_rules_args(rows, cols, count); if (rc) return rc;
  if (count == 0) return YY_OK;
  int cells = rows * cols, W = words_for_cells(cells);
  unsigned grid = (unsigned)((count + kRulesBlock - 1) / kRulesBlock);
  const int sq = sq_side(rows, cols, rule_flags);
  if (sq == 8) legal_mask_sq_kernel<8><<<grid, kRulesBlock, 0, (cudaStream_t)stream>>>(black, white, players, out_mask, count);
  else if (sq == 6) legal_mask_sq_kernel<6><<<grid, kRulesBlock, 0, (cudaStream_t)stream>>>(black, white, players, out_mask, count);
  else YY_DISPATCH_NW(cells, legal_mask_kernel<NW><<<grid, kRulesBlock, 0, (cudaStream_t)stream>>>(
      make_geo<NW>(rows, cols, rule_flags), W, black, white, players, out_mask, count));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_step(int rows, int cols, uint32_t rule_flags, uint64_t* black, uint64_t* white, int8_t* players,
            const int32_t* actions, int64_t count, void* stream) {
  int rc = check_rules_args(rows, cols, count); if (rc) return rc;
  if (count == 0) return YY_OK;
  int cells = rows * cols, W = words_for_cells(cells);
  unsigned grid = (unsigned)((count + kRulesBlock - 1) / kRulesBlock);
  const int sq = sq_side(rows, cols, rule_flags);
  if (sq == 8) step_sq_kernel<8><<<grid, kRulesBlock, 0, (cudaStream_t)stream>>>(black, white, players, actions, count);
  else if (sq == 6) step_sq_kernel<6><<<grid, kRulesBlock, 0, (cudaStream_t)stream>>>(black, white, players, actions, count);
  else YY_DISPATCH_NW(cells, step_kernel<NW><<<grid, kRulesBlock, 0, (cudaStream_t)stream>>>(
      make_geo<NW>(rows, cols, rule_flags), W, black, white, players, actions, count));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_ended(int rows, int cols, uint32_t rule_flags, const uint64_t* black, const uint64_t* white,
             const int8_t* players, int8_t* out_result, int64_t count, void* stream) {
  int rc = check_rules_args(rows, cols, count); if (rc) return rc;
  if (count == 0) return YY_OK;
  int cells = rows * cols, W = words_for_cells(cells);
  unsigned grid = (unsigned)((count + kRulesBlock - 1) / kRulesBlock);
  const int sq = sq_side(rows, cols, rule_flags);
  if (sq == 8) ended_sq_kernel<8><<<grid, kRulesBlock, 0, (cudaStream_t)stream>>>(black, white, players, out_result, count);
  else if (sq == 6) ended_sq_kernel<6><<<grid, kRulesBlock, 0, (cudaStream_t)stream>>>(black, white, players, out_result, count);
  else YY_DISPATCH_NW(cells, ended_kernel<NW><<<grid, kRulesBlock, 0, (cudaStream_t)stream>>>(
      make_geo<NW>(rows, cols, rule_flags), W, black, white, players, out_result, count));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_env_step(int rows, int cols, uint32_t rule_flags, uint64_t* black, uint64_t* white, int8_t* players,
                const int32_t* actions, uint64_t* out_mask, int8_t* out_result, int64_t count, void* stream) {
  int rc = check_rules_args(rows, cols, count); if (rc) return rc;
  if (count == 0) return YY_OK;
  int cells = rows * cols, W = words_for_cells(cells);
  // Blocks counting-sort their boards by stone count so that a warp's lanes run fills of similar length.  64 boards per
  // block (128 threads) spread a batch more evenly over the 148 SMs than 128 (65,536 boards: 6.9 blocks per SM instead of
  // 3.5, i.e. the fullest SM holds 1 % more than the average instead of 15 %): 6.5 against 7.1 us per launch, and still
  // 4 % ahead at 1 M boards.
  constexpr int kBlock = 128;
#define YY_ENV_LAUNCH(NWV, SIDEV)                                                                            \
  env_step_kernel<NWV, kBlock, true, SIDEV><<<(unsigned)((count + kBlock / 2 - 1) / (kBlock / 2)), kBlock, 0, (cudaStream_t)stream>>>( \
      make_geo<NWV>(rows, cols, rule_flags), W, black, white, players, actions, out_mask, out_result, count)
  const int side = rows == cols ? rows : 0;
  if (sq_side(rows, cols, rule_flags)) {
    constexpr int kSqBlock = 128;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kSqBlock);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const long long kChunk = 1ll << 30;        // boards per launch: the kernel indexes its lanes with 32 bits
    for (long long off = 0; off < count; off += kChunk) {
      const long long n = count - off < kChunk ? count - off : kChunk;
      cfg.gridDim = dim3((unsigned)((2 * n + kSqBlock - 1) / kSqBlock));
      uint64_t *bk = black + off, *wh = white + off, *om = out_mask + off;
      int8_t *pl = players + off, *res = out_result + off;
      const int32_t* ac = actions + off;
      if (side == 8) YY_CUDA_OK(cudaLaunchKernelEx(&cfg, env_step_sq_kernel<8, kSqBlock>, bk, wh, pl, ac, om, res, n));
      else YY_CUDA_OK(cudaLaunchKernelEx(&cfg, env_step_sq_kernel<6, kSqBlock>, bk, wh, pl, ac, om, res, n));
      if (off) count_launch();
    }
  }
  else if (side == 8) YY_ENV_LAUNCH(1, 8);
  else if (side == 6) YY_ENV_LAUNCH(1, 6);
  else if (side == 16) YY_ENV_LAUNCH(4, 16);
  else YY_DISPATCH_NW(cells, YY_ENV_LAUNCH(NW, 0));
#undef YY_ENV_LAUNCH
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_pack_boards(int rows, int cols, const int8_t* boards, uint64_t* black, uint64_t* white, int64_t count, void* stream) {
  int rc = check_rules_args(rows, cols, count); if (rc) return rc;
  if (count == 0) return YY_OK;
  const int cells = rows * cols, W = words_for_cells(cells);
  pack_boards_kernel<<<(unsigned)((count * W + 255) / 256), 256, 0, (cudaStream_t)stream>>>(boards, cells, W, black, white, count);
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_unpack_boards(int rows, int cols, const uint64_t* black, const uint64_t* white, const uint64_t* mask, int8_t* out_boards,
                     uint8_t* out_mask, int64_t count, void* stream) {
  int rc = check_rules_args(rows, cols, count); if (rc) return rc;
  if (count == 0 || (!out_boards && !out_mask)) return YY_OK;
  if ((out_boards && (!black || !white)) || (out_mask && !mask)) return set_error(YY_ERR_INVALID, "unpack: missing source words");
  const int cells = rows * cols, W = words_for_cells(cells);
  unpack_boards_kernel<<<(unsigned)((count * cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(black, white, mask, cells, W, out_boards, out_mask, count);
  YY_LAUNCH_CHECK();
  return YY_OK;
}

int yy_random_playout(int rows, int cols, uint32_t rule_flags, uint64_t seed, const int32_t* plies, uint64_t* black,
                      uint64_t* white, int8_t* players, int64_t count, void* stream) {
  int rc = check_rules_args(rows, cols, count); if (rc) return rc;
  if (count == 0) return YY_OK;
  int cells = rows * cols, W = words_for_cells(cells);
  unsigned grid = (unsigned)((count + kRulesBlock - 1) / kRulesBlock);
  YY_DISPATCH_NW(cells, random_playout_kernel<NW><<<grid, kRulesBlock, 0, (cudaStream_t)stream>>>(
      make_geo<NW>(rows, cols, rule_flags), W, seed, plies, black, white, players, count));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

}  // extern "C"
