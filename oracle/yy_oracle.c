/*
 * oracle/yy_oracle.c -- CPU restatement of the Yin-Yang self-play hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may include, link
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Every function cites the reference file:line (relative to the upstream
 * repository root) whose behaviour it restates.  The restatement is plain C on
 * int8 boards (0 empty, +1 black, -1 white), deliberately *not* bitboards, so
 * that it is an independent check of the CUDA bitboard kernels.
 *
 * Parity pin: tests/golden/ (generated from the live Python reference by
 * tests/golden/make_golden.py) -- see tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define YO_MAX_CELLS 1024

/* rule_flags bit 0: also apply the "no completed single-colour row/column"
 * rule of src/gui/static/js/yin_yang_game.js:338-384 (absent from Python). */
#define YO_RULE_ROWCOL 1

/* ---------------------------------------------------------------- rules --- */

/* src/yin_yang/yin_yang_logic.py:58-94 (_check_connectivity): every cell of
 * colour `piece` must be reachable from the first one in row-major order. */
static int yo_connected(const int8_t *b, int n, int m, int piece) {
  int total = 0, first = -1;
  for (int i = 0; i < n * m; ++i)
    if (b[i] == piece) { if (first < 0) first = i; ++total; }
  if (total == 0) return 1;
  static _Thread_local int stack[YO_MAX_CELLS];
  static _Thread_local uint8_t seen[YO_MAX_CELLS];
  memset(seen, 0, (size_t)(n * m));
  int sp = 0, reached = 1;
  stack[sp++] = first; seen[first] = 1;
  while (sp) {
    int c = stack[--sp], x = c / m, y = c % m;
    const int dx[4] = {0, 1, 0, -1}, dy[4] = {1, 0, -1, 0};
    for (int k = 0; k < 4; ++k) {
      int nx = x + dx[k], ny = y + dy[k];
      if (nx < 0 || nx >= n || ny < 0 || ny >= m) continue;
      int d = nx * m + ny;
      if (b[d] == piece && !seen[d]) { seen[d] = 1; ++reached; stack[sp++] = d; }
    }
  }
  return reached == total;
}

/* src/yin_yang/yin_yang_logic.py:96-109 (_check_2x2_constraint): no 2x2 window
 * anywhere on the board holds four equal non-empty cells (either colour). */
static int yo_no_2x2(const int8_t *b, int n, int m) {
  for (int i = 0; i + 1 < n; ++i)
    for (int j = 0; j + 1 < m; ++j) {
      int8_t v = b[i * m + j];
      if (v != 0 && b[i * m + j + 1] == v && b[(i + 1) * m + j] == v && b[(i + 1) * m + j + 1] == v)
        return 0;
    }
  return 1;
}

/* src/gui/static/js/yin_yang_game.js:338-384 (checkRowColumnConstraint): a row
 * or column with no empty cell and only one colour is a violation. */
static int yo_rowcol_ok(const int8_t *b, int n, int m) {
  for (int i = 0; i < n; ++i) {
    int hb = 0, hw = 0, he = 0;
    for (int j = 0; j < m; ++j) { int8_t v = b[i * m + j]; hb |= v == 1; hw |= v == -1; he |= v == 0; }
    if (!he && (!hb || !hw)) return 0;
  }
  for (int j = 0; j < m; ++j) {
    int hb = 0, hw = 0, he = 0;
    for (int i = 0; i < n; ++i) { int8_t v = b[i * m + j]; hb |= v == 1; hw |= v == -1; he |= v == 0; }
    if (!he && (!hb || !hw)) return 0;
  }
  return 1;
}

/* src/yin_yang/yin_yang_logic.py:31-56 (is_valid_move): in bounds, empty, and
 * after a trial placement: connectivity of `piece`, then the global 2x2 test
 * (+ optional JS row/column test, yin_yang_game.js:187-232 order). */
int yo_is_valid_move(const int8_t *board, int n, int m, int x, int y, int piece, int rule_flags) {
  if (x < 0 || x >= n || y < 0 || y >= m) return 0;
  if (board[x * m + y] != 0) return 0;
  int8_t tmp[YO_MAX_CELLS];
  memcpy(tmp, board, (size_t)(n * m));
  tmp[x * m + y] = (int8_t)piece;
  if (!yo_connected(tmp, n, m, piece)) return 0;
  if (!yo_no_2x2(tmp, n, m)) return 0;
  if ((rule_flags & YO_RULE_ROWCOL) && !yo_rowcol_ok(tmp, n, m)) return 0;
  return 1;
}

/* src/yin_yang/yin_yang_logic.py:111-120 + yin_yang_game.py:60-78
 * (get_valid_moves / getValidMoves): mask[a] = 1 iff action a = x*m+y legal. */
int yo_legal_mask(const int8_t *board, int n, int m, int player, int rule_flags, uint8_t *mask) {
  int piece = player == 1 ? 1 : -1, cnt = 0;
  for (int x = 0; x < n; ++x)
    for (int y = 0; y < m; ++y) {
      int ok = yo_is_valid_move(board, n, m, x, y, piece, rule_flags);
      mask[x * m + y] = (uint8_t)ok; cnt += ok;
    }
  return cnt;
}

/* src/yin_yang/yin_yang_logic.py:122-128 (has_valid_move). */
int yo_has_valid_move(const int8_t *board, int n, int m, int piece, int rule_flags) {
  for (int x = 0; x < n; ++x)
    for (int y = 0; y < m; ++y)
      if (yo_is_valid_move(board, n, m, x, y, piece, rule_flags)) return 1;
  return 0;
}

/* src/yin_yang/yin_yang_game.py:39-58 + yin_yang_logic.py:24-29 (getNextState /
 * place_piece): place if legal, otherwise silently leave the board unchanged;
 * the turn passes either way.  Value semantics: `board` is updated in place by
 * the caller's choice (callers pass a copy). */
int yo_next_state(int8_t *board, int n, int m, int player, int action, int rule_flags) {
  int piece = player == 1 ? 1 : -1;
  int x = action / m, y = action % m;
  if (action >= 0 && yo_is_valid_move(board, n, m, x, y, piece, rule_flags)) board[x * m + y] = (int8_t)piece;
  return -player;
}

/* src/yin_yang/yin_yang_game.py:80-110 (getGameEnded) + yin_yang_logic.py:130-134:
 * 0 ongoing; terminal iff neither colour can move; then +1/-1 from `player`'s
 * perspective by piece count, draw = +0.0001. */
double yo_game_ended(const int8_t *board, int n, int m, int player, int rule_flags) {
  int pp = player == 1 ? 1 : -1;
  if (yo_has_valid_move(board, n, m, pp, rule_flags)) return 0.0;
  if (yo_has_valid_move(board, n, m, -pp, rule_flags)) return 0.0;
  int bc = 0, wc = 0;
  for (int i = 0; i < n * m; ++i) { bc += board[i] == 1; wc += board[i] == -1; }
  if (bc > wc) return player == 1 ? 1.0 : -1.0;
  if (wc > bc) return player == -1 ? 1.0 : -1.0;
  return 0.0001;
}

/* Batched helpers for numpy callers: boards int8[count][n*m]. */
void yo_legal_mask_batch(const int8_t *boards, const int8_t *players, int64_t count, int n, int m,
                         int rule_flags, uint8_t *masks) {
  for (int64_t i = 0; i < count; ++i)
    yo_legal_mask(boards + i * n * m, n, m, players[i], rule_flags, masks + i * n * m);
}
void yo_next_state_batch(int8_t *boards, int8_t *players, const int32_t *actions, int64_t count, int n, int m,
                         int rule_flags) {
  for (int64_t i = 0; i < count; ++i)
    players[i] = (int8_t)yo_next_state(boards + i * n * m, n, m, players[i], actions[i], rule_flags);
}
void yo_game_ended_batch(const int8_t *boards, const int8_t *players, int64_t count, int n, int m, int rule_flags,
                         double *out) {
  for (int64_t i = 0; i < count; ++i) out[i] = yo_game_ended(boards + i * n * m, n, m, players[i], rule_flags);
}

/* One "env step" of BASELINE.md: mask for the side to move, apply the action,
 * terminal/score of the successor from the next player's perspective. */
void yo_env_step_batch(int8_t *boards, int8_t *players, const int32_t *actions, int64_t count, int n, int m,
                       int rule_flags, uint8_t *masks, double *results) {
  for (int64_t i = 0; i < count; ++i) {
    int8_t *b = boards + i * n * m;
    yo_legal_mask(b, n, m, players[i], rule_flags, masks + i * n * m);
    players[i] = (int8_t)yo_next_state(b, n, m, players[i], actions[i], rule_flags);
    results[i] = yo_game_ended(b, n, m, players[i], rule_flags);
  }
}

/* ------------------------------------------------ deterministic evaluator --- */
/* Hash-stub evaluator shared (by specification, not by code) with the CUDA
 * engine's deterministic-prior mode.  Priors are dyadic k/2^16 (k in 1..4096),
 * value is j/2^16 in [-1,1): exactly representable in float32, so the search
 * arithmetic is identical under numpy-1 float64 and numpy-2 float32 rules.
 * Like the reference network (neural_network.py:156-196: no side-to-move
 * plane) it is a function of the board only. */
static inline uint64_t yo_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
uint64_t yo_stub_key(const int8_t *board, int cells) {
  int words = (cells + 63) / 64;
  uint64_t key = 0x243F6A8885A308D3ull;
  for (int pass = 0; pass < 2; ++pass) {
    int8_t want = pass == 0 ? 1 : -1;
    for (int w = 0; w < words; ++w) {
      uint64_t bits = 0;
      for (int k = 0; k < 64 && w * 64 + k < cells; ++k)
        if (board[w * 64 + k] == want) bits |= 1ull << k;
      key = yo_mix64(key ^ bits);
    }
  }
  return key;
}
void yo_stub_predict(const int8_t *board, int cells, float *policy, float *value) {
  uint64_t key = yo_stub_key(board, cells);
  for (int a = 0; a < cells; ++a)
    policy[a] = (float)(1u + (uint32_t)(yo_mix64(key + 0xD1B54A32D192ED03ull * (uint64_t)(a + 1)) >> 52)) / 65536.0f;
  *value = (float)((int32_t)(yo_mix64(key ^ 0xA0761D6478BD642Full) >> 47) - 65536) / 65536.0f;
}

/* ----------------------------------------------------------------- MCTS --- */

typedef void (*yo_eval_fn)(const int8_t *board, int n, int m, float *policy, float *value, void *ctx);

typedef struct {
  int parent, action;        /* mcts.py:31-34 */
  int visits;                /* mcts.py:38 */
  float value_sum;           /* mcts.py:39: float32 after the first update (numpy>=2, NEP 50) */
  float prior;               /* mcts.py:40 */
  int first_child, n_children;
  int player;                /* mcts.py:44 */
  int is_terminal;           /* mcts.py:46 */
  double terminal_value;     /* mcts.py:47 (Python int/float) */
  int has_state;
  int depth;
} yo_node;

typedef struct {
  int n, m, rule_flags;
  float cpuct;
  yo_eval_fn eval; void *eval_ctx;
  yo_node *nodes; int n_nodes, cap_nodes;
  int8_t *boards; /* cap_nodes * n*m, valid where has_state */
  int64_t n_evals;
} yo_tree;

static int yo_new_node(yo_tree *t, int parent, int action, float prior) {
  if (t->n_nodes == t->cap_nodes) {
    t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
    t->nodes = (yo_node *)realloc(t->nodes, sizeof(yo_node) * (size_t)t->cap_nodes);
    t->boards = (int8_t *)realloc(t->boards, (size_t)t->cap_nodes * (size_t)(t->n * t->m));
  }
  yo_node *nd = &t->nodes[t->n_nodes];
  memset(nd, 0, sizeof(*nd));
  nd->parent = parent; nd->action = action; nd->prior = prior; nd->first_child = -1;
  nd->depth = parent >= 0 ? t->nodes[parent].depth + 1 : 0;
  return t->n_nodes++;
}

static void yo_do_eval(yo_tree *t, const int8_t *board, float *policy, float *value) {
  ++t->n_evals;
  if (t->eval) t->eval(board, t->n, t->m, policy, value, t->eval_ctx);
  else yo_stub_predict(board, t->n * t->m, policy, value);
}

/* mcts.py:50-91 (Node.expand): terminal test first, then one child per legal
 * action in ascending action order with prior = policy[a] (raw, unmasked,
 * un-normalised: mcts.py:77-78). */
static void yo_expand(yo_tree *t, int id, const int8_t *board, int player, const float *policy) {
  int cells = t->n * t->m;
  int8_t *dst = t->boards + (size_t)id * cells;
  if (dst != board) memcpy(dst, board, (size_t)cells);
  t->nodes[id].has_state = 1; t->nodes[id].player = player;
  double res = yo_game_ended(dst, t->n, t->m, player, t->rule_flags);
  if (res != 0.0) { t->nodes[id].is_terminal = 1; t->nodes[id].terminal_value = res; return; }
  uint8_t mask[YO_MAX_CELLS];
  yo_legal_mask(dst, t->n, t->m, player, t->rule_flags, mask);
  if (t->nodes[id].n_children > 0) return; /* re-expansion of a child-less node never reaches here with children */
  int first = -1, cnt = 0;
  for (int a = 0; a < cells; ++a)
    if (mask[a]) {
      int c = yo_new_node(t, id, a, policy[a]);
      if (first < 0) first = c;
      ++cnt;
    }
  t->nodes[id].first_child = first; t->nodes[id].n_children = cnt;
}

/* mcts.py:97-145 (Node.select_child), numpy>=2 float32 op order:
 *   S = sum child visits (int);  t1 = f32(c*P);  t2 = f32(t1 * f32(sqrt_f64(S)));
 *   u = f32(t2 / f32(1+N));  q = N>0 ? f32(W / f32(N)) : 0;  ucb = f32(q+u);
 * strict '>' in ascending action order => lowest action among ties. */
static int yo_select_child(const yo_tree *t, int id) {
  const yo_node *nd = &t->nodes[id];
  long sum = 0;
  for (int k = 0; k < nd->n_children; ++k) sum += t->nodes[nd->first_child + k].visits;
  volatile float sq = (float)sqrt((double)sum);
  int best = -1; float best_ucb = -INFINITY;
  for (int k = 0; k < nd->n_children; ++k) {
    const yo_node *ch = &t->nodes[nd->first_child + k];
    volatile float t1 = t->cpuct * ch->prior;
    volatile float t2 = t1 * sq;
    volatile float u = t2 / (float)(1 + ch->visits);
    volatile float q = 0.0f;
    if (ch->visits > 0) q = ch->value_sum / (float)ch->visits;
    volatile float ucb = q + u;
    if (ucb > best_ucb) { best_ucb = ucb; best = nd->first_child + k; }
  }
  return best;
}

/* mcts.py:345-414 (_simulate) + :147-156 (Node.update). */
static void yo_simulate(yo_tree *t, int root) {
  int cells = t->n * t->m;
  int path[YO_MAX_CELLS + 2]; int plen = 0;
  int cur = root; path[plen++] = cur;
  while ((t->nodes[cur].n_children > 0 || t->nodes[cur].is_terminal) && !t->nodes[cur].is_terminal) {
    cur = yo_select_child(t, cur);
    path[plen++] = cur;
  }
  float value; /* every update lands in a float32 accumulator */
  float policy[YO_MAX_CELLS];
  if (t->nodes[cur].is_terminal) {
    value = (float)t->nodes[cur].terminal_value; /* python scalar, weak-promoted into f32 add */
  } else if (plen < 2) { /* mcts.py:371-382: root without children re-evaluated */
    int8_t tmp[YO_MAX_CELLS]; memcpy(tmp, t->boards + (size_t)root * cells, (size_t)cells);
    yo_do_eval(t, tmp, policy, &value);
    yo_expand(t, root, tmp, t->nodes[root].player, policy);
  } else { /* mcts.py:384-397 */
    int par = path[plen - 2];
    int8_t nb[YO_MAX_CELLS]; memcpy(nb, t->boards + (size_t)par * cells, (size_t)cells);
    int np_ = yo_next_state(nb, t->n, t->m, t->nodes[par].player, t->nodes[cur].action, t->rule_flags);
    yo_do_eval(t, nb, policy, &value);
    yo_expand(t, cur, nb, np_, policy);
  }
  int leaf_player = t->nodes[path[plen - 1]].player;
  for (int i = plen - 1; i >= 0; --i) { /* mcts.py:406-412 */
    yo_node *nd = &t->nodes[path[i]];
    float v = (i != plen - 1 && nd->player != leaf_player) ? -value : value;
    nd->visits += 1;
    volatile float s = nd->value_sum + v;
    nd->value_sum = s;
  }
}

/* mcts.py:275-343 (MCTS.search), num_threads == 1.  `noise` (may be NULL) holds
 * one float64 Dirichlet sample per legal root action in ascending action order
 * and is mixed as mcts.py:309-311 does under numpy>=2:
 *   policy[a] = f32( f64(f32(f32(1-eps) * p)) + eps * noise_i ).
 * Outputs: counts[A] (root.get_children_visit_counts, mcts.py:168-181),
 * child_w[A] (value_sum of each root child), returns number of evaluator calls. */
int64_t yo_mcts_search(const int8_t *board, int n, int m, int player, int rule_flags, int num_sims, float cpuct,
                       const double *noise, double eps, yo_eval_fn eval, void *eval_ctx, int32_t *counts,
                       float *child_w, int32_t *out_stats /* [n_nodes, max_depth, root_visits] or NULL */) {
  yo_tree t; memset(&t, 0, sizeof(t));
  t.n = n; t.m = m; t.rule_flags = rule_flags; t.cpuct = cpuct; t.eval = eval; t.eval_ctx = eval_ctx;
  int cells = n * m;
  int root = yo_new_node(&t, -1, -1, 0.0f);
  float policy[YO_MAX_CELLS], value;
  yo_do_eval(&t, board, policy, &value);
  if (noise) {
    uint8_t mask[YO_MAX_CELLS];
    yo_legal_mask(board, n, m, player, rule_flags, mask);
    int k = 0;
    for (int a = 0; a < cells; ++a)
      if (mask[a]) {
        volatile float keep = (float)(1.0 - eps) * policy[a];
        policy[a] = (float)((double)keep + eps * noise[k++]);
      }
  }
  yo_expand(&t, root, board, player, policy);
  for (int s = 0; s < num_sims; ++s) yo_simulate(&t, root);
  for (int a = 0; a < cells; ++a) { counts[a] = 0; if (child_w) child_w[a] = 0.0f; }
  const yo_node *r = &t.nodes[root];
  for (int k = 0; k < r->n_children; ++k) {
    const yo_node *ch = &t.nodes[r->first_child + k];
    counts[ch->action] = ch->visits;
    if (child_w) child_w[ch->action] = ch->value_sum;
  }
  if (out_stats) {
    int md = 0;
    for (int i = 0; i < t.n_nodes; ++i) if (t.nodes[i].has_state && t.nodes[i].depth > md) md = t.nodes[i].depth;
    out_stats[0] = t.n_nodes; out_stats[1] = md; out_stats[2] = r->visits;
  }
  int64_t ev = t.n_evals;
  free(t.nodes); free(t.boards);
  return ev;
}
