// yy_rules_sq.cuh -- Yin-Yang rules for square boards that fit ONE 64-bit word (side <= 8), the board side a
// compile-time constant.  Same predicates as yy_rules.cuh (is_valid_move src/yin_yang/yin_yang_logic.py:31-56,
// _check_connectivity :58-94, _check_2x2_constraint :96-109), written for the shortest instruction chain:
//
//  * connectivity by LINE fills instead of one-cell dilation steps.  A round extends the filled set along whole
//    horizontal runs and then along whole vertical runs of the colour:
//      - east: the run above a seed is cleared by the carry of ONE addition (x + seeds; the last column is kept
//        out of the addend so that a carry never enters the next row, and is patched from the carry it receives);
//      - west: the same on the board rotated by 180 degrees (one bit reversal each way);
//      - north / south: Kogge-Stone occluded fills, 3 doubling steps, their propagators computed once per colour.
//    Random-play 8x8 boards need 2.6 rounds on average (at most 9) where the one-cell flood needs 9.5 steps
//    (at most 24); the fixed point is the same set, so legality is unchanged bit for bit.
//  * the 2x2 rule from the horizontal pairs H = x & east(x): has_2x2 = H & (H >> side); a cell completes a 2x2 when a
//    pair lies in the row above or below next to it and the cell between them is taken -- 18 instructions instead of 36.
//
// __host__ __device__ so that tests/_host can check the algebra on the CPU box; the product only runs it in kernels.
#pragma once
#include "yy_rules.cuh"

namespace yy {

YY_HD uint64_t brev64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __brevll(v);
#else
  v = ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
  v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
  v = ((v >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((v & 0x0F0F0F0F0F0F0F0Full) << 4);
  return __builtin_bswap64(v);
#endif
}

template <int SIDE>
struct Sq {
  static_assert(SIDE >= 2 && SIDE <= 8, "one 64-bit word per colour");
  static constexpr int S = SIDE, CELLS = SIDE * SIDE;
  static constexpr uint64_t FULL = CELLS == 64 ? ~0ull : ((1ull << (CELLS & 63)) - 1);
  static constexpr uint64_t column(int c) {
    uint64_t m = 0;
    for (int r = 0; r < SIDE; ++r) m |= 1ull << (r * SIDE + c);
    return m;
  }
  static constexpr uint64_t COL0 = column(0), COLL = column(SIDE - 1);
  static constexpr uint64_t NOT0 = FULL & ~COL0, NOTL = FULL & ~COLL;

  // the board turned by 180 degrees: cell (r, c) -> (S-1-r, S-1-c); east on the turned board is west on the board
  static YY_HD uint64_t turn(uint64_t v) { return brev64(v) >> (64 - CELLS); }
  static YY_HD uint64_t dilate4(uint64_t x) { return ((x << 1) & NOT0) | ((x >> 1) & NOTL) | ((x << S) & FULL) | (x >> S); }

  // _check_2x2_constraint for one colour and the cells whose occupation by that colour would complete a 2x2 block
  static YY_HD void blocks(uint64_t x, bool& has, uint64_t& completes) {
    const uint64_t x1 = (x >> 1) & NOTL;          // cells whose east neighbour is x
    const uint64_t H = x & x1;                    // horizontal pairs, marked at their west cell
    const uint64_t below = H >> S, above = H << S;  // a pair in the next / previous row, same columns
    has = (H & below) != 0;
    const uint64_t ab = above | below;
    completes = (ab & x1) | ((ab & x) << 1);      // the cell west of / east of the third stone
  }

  // runs of p (last column split off: pn = p & NOTL, pc = p & COLL) reached from the seeds f towards higher columns
  static YY_HD uint64_t east(uint64_t f, uint64_t pn, uint64_t pc) {
    const uint64_t s = pn + (f & NOTL);
    return (pn & ~s) | (pc & s);
  }

  struct Lines {                                  // what a fill needs of the colour x, computed once
    uint64_t x, xn, xc, tn, tc, n1, n2, s1, s2;
  };
  static YY_HD Lines lines(uint64_t x) {
    Lines l;
    const uint64_t t = turn(x);
    l.x = x; l.xn = x & NOTL; l.xc = x & COLL; l.tn = t & NOTL; l.tc = t & COLL;
    l.n1 = x & (x >> S); l.n2 = l.n1 & (l.n1 >> (2 * S));
    l.s1 = x & (x << S); l.s2 = l.s1 & (l.s1 << (2 * S));
    return l;
  }
  // one round: whole horizontal runs through f, then whole vertical runs through those
  static YY_HD uint64_t round(const Lines& l, uint64_t f) {
    const uint64_t h = f | east(f, l.xn, l.xc) | turn(east(turn(f), l.tn, l.tc));
    uint64_t a = h, b = h;
    a |= l.x & (a >> S); a |= l.n1 & (a >> (2 * S)); a |= l.n2 & (a >> (4 * S));
    b |= l.x & (b << S); b |= l.s1 & (b << (2 * S)); b |= l.s2 & (b << (4 * S));
    return a | b;
  }

  // cells adjacent to EVERY 4-connected component of x (all cells if x is empty); components are only examined while
  // some cell of `wanted` can still qualify.  Equals touches_all_components() of yy_rules.cuh on wanted.
  static YY_HD uint64_t touches_all(uint64_t x, uint64_t wanted) {
    if (!x) return FULL;
    const Lines l = lines(x);
    uint64_t acc = FULL, rem = x;
    do {
      uint64_t f = rem & (0 - rem), v;
      for (;;) {
        v = round(l, f);
        if (v == rem || v == f) break;            // the component is all that was left / nothing new this round
        f = v;
      }
      acc &= dilate4(v);
      rem &= ~v;
    } while (rem && (acc & wanted));
    return acc;
  }

  // get_valid_moves (yin_yang_logic.py:111-120) for the colour with stones p against o
  static YY_HD uint64_t legal(uint64_t p, uint64_t o) {
    bool hp, ho; uint64_t cp, co;
    blocks(p, hp, cp); blocks(o, ho, co);
    if (hp || ho) return 0;
    const uint64_t cand = FULL & ~(p | o) & ~cp;
    if (!cand) return 0;
    return cand & touches_all(p, cand);
  }
};

}  // namespace yy
