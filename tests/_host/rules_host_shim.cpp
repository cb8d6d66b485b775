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
