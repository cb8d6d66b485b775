// Test-only: compiles the product's bitboard rules header (yy_rules.cuh) for the HOST with g++ so
// that the CPU suite can check the bitboard algorithm against the oracle without a GPU.
// Never linked into the product library.
#include <cstring>
#include "../../yinyang-game-alphazero_b200/csrc/yy_rules.cuh"
using namespace yy;

template <int NW>
static void run(int rows, int cols, unsigned flags, const uint64_t* black, const uint64_t* white, const int8_t* players,
                const int32_t* actions, long count, uint64_t* mask_out, uint64_t* nb_out, uint64_t* nw_out,
                int8_t* np_out, int8_t* ended_out, float* stub_out) {
  Geo<NW> g = make_geo<NW>(rows, cols, flags);
  int W = (rows * cols + 63) / 64;
  for (long i = 0; i < count; ++i) {
    BB<NW> b = bb_zero<NW>(), w = bb_zero<NW>();
    for (int k = 0; k < W; ++k) { b.w[k] = black[i * W + k]; w.w[k] = white[i * W + k]; }
    b = b & g.full; w = w & g.full;
    BB<NW> lm = legal_for(g, b, w, players[i]);
    for (int k = 0; k < W; ++k) mask_out[i * W + k] = lm.w[k];
    ended_out[i] = (int8_t)ended_code(g, b, w, players[i]);
    if (stub_out) {
      uint64_t key = stub_key(g, b, w);
      for (int a = 0; a < g.cells; ++a) stub_out[i * (g.cells + 1) + a] = stub_prior(key, a);
      stub_out[i * (g.cells + 1) + g.cells] = stub_value(key);
    }
    apply_action(g, b, w, players[i], actions[i]);
    for (int k = 0; k < W; ++k) { nb_out[i * W + k] = b.w[k]; nw_out[i * W + k] = w.w[k]; }
    np_out[i] = (int8_t)-players[i];
  }
}

extern "C" int yyh_rules(int rows, int cols, unsigned flags, const uint64_t* black, const uint64_t* white,
                         const int8_t* players, const int32_t* actions, long count, uint64_t* mask_out,
                         uint64_t* nb_out, uint64_t* nw_out, int8_t* np_out, int8_t* ended_out, float* stub_out) {
  int cells = rows * cols;
  if (cells <= 64) run<1>(rows, cols, flags, black, white, players, actions, count, mask_out, nb_out, nw_out, np_out, ended_out, stub_out);
  else if (cells <= 128) run<2>(rows, cols, flags, black, white, players, actions, count, mask_out, nb_out, nw_out, np_out, ended_out, stub_out);
  else if (cells <= 256) run<4>(rows, cols, flags, black, white, players, actions, count, mask_out, nb_out, nw_out, np_out, ended_out, stub_out);
  else return -1;
  return 0;
}

// The line-fill rules for square one-word boards (yy_rules_sq.cuh), same outputs as yyh_rules (no stub evaluator).
#include "../../yinyang-game-alphazero_b200/csrc/yy_rules_sq.cuh"
template <int SIDE>
static void run_sq(const uint64_t* black, const uint64_t* white, const int8_t* players, const int32_t* actions, long count,
                   uint64_t* mask_out, uint64_t* nb_out, uint64_t* nw_out, int8_t* np_out, int8_t* ended_out) {
  typedef Sq<SIDE> Q;
  for (long i = 0; i < count; ++i) {
    uint64_t b = black[i] & Q::FULL, w = white[i] & Q::FULL;
    const bool mb = players[i] == 1;
    const uint64_t p = mb ? b : w, o = mb ? w : b;
    const uint64_t lm = Q::legal(p, o);
    mask_out[i] = lm;
    int code = 0;
    if (!lm && !Q::legal(o, p)) { int mc = popc64(p), oc = popc64(o); code = mc > oc ? 1 : (oc > mc ? -1 : 2); }
    ended_out[i] = (int8_t)code;
    const int a = actions[i];
    if (a >= 0 && a < Q::CELLS && ((lm >> a) & 1)) { if (mb) b |= 1ull << a; else w |= 1ull << a; }
    nb_out[i] = b; nw_out[i] = w; np_out[i] = (int8_t)-players[i];
  }
}
extern "C" int yyh_rules_sq(int side, const uint64_t* black, const uint64_t* white, const int8_t* players,
                            const int32_t* actions, long count, uint64_t* mask_out, uint64_t* nb_out, uint64_t* nw_out,
                            int8_t* np_out, int8_t* ended_out) {
  switch (side) {
    case 3: run_sq<3>(black, white, players, actions, count, mask_out, nb_out, nw_out, np_out, ended_out); return 0;
    case 4: run_sq<4>(black, white, players, actions, count, mask_out, nb_out, nw_out, np_out, ended_out); return 0;
    case 5: run_sq<5>(black, white, players, actions, count, mask_out, nb_out, nw_out, np_out, ended_out); return 0;
    case 6: run_sq<6>(black, white, players, actions, count, mask_out, nb_out, nw_out, np_out, ended_out); return 0;
    case 7: run_sq<7>(black, white, players, actions, count, mask_out, nb_out, nw_out, np_out, ended_out); return 0;
    case 8: run_sq<8>(black, white, players, actions, count, mask_out, nb_out, nw_out, np_out, ended_out); return 0;
  }
  return -1;
}
