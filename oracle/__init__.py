"""CPU oracle for the Yin-Yang self-play hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.  The
product package (``yinyang-game-alphazero_b200/``) never imports it.

Contents
  yy_oracle.c   plain-C restatement (int8 boards, BFS) of the reference rules
                (src/yin_yang/yin_yang_logic.py, yin_yang_game.py) and of the
                sequential MCTS (src/yin_yang/ai/mcts.py) in numpy>=2 float32
                op order, plus the deterministic hash-stub evaluator.
  port.py       Python/numpy/torch port of the reference's CPU self-play path
                (same algorithmic structure and cost profile: per-cell BFS
                legality, object tree, batch-1 fp32 torch forward).  This is what
                ``cpu_baseline`` / ``--impl reference`` time (kind "port").

Parity pin: tests/golden/*.npz were produced by the *unmodified* Python reference
(imported from /root/reference in the build container) by
tests/golden/make_golden.py; tests/test_oracle_golden.py checks this oracle
against them bit-for-bit.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libyy_oracle.so")
RULE_ROWCOL = 1


def build(force: bool = False) -> str:
    """Compile yy_oracle.c with gcc (seconds).  Strict IEEE float32: no FMA contraction."""
    src = os.path.join(_HERE, "yy_oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-o", _SO, src, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


_lib = None
EVAL_FN = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_int8), ctypes.c_int, ctypes.c_int,
                           ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.c_void_p)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        i8p, u8p, i32p = (ctypes.POINTER(ctypes.c_int8), ctypes.POINTER(ctypes.c_uint8),
                          ctypes.POINTER(ctypes.c_int32))
        f32p, f64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
        L.yo_legal_mask_batch.argtypes = [i8p, i8p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p]
        L.yo_next_state_batch.argtypes = [i8p, i8p, i32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.yo_game_ended_batch.argtypes = [i8p, i8p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, f64p]
        L.yo_env_step_batch.argtypes = [i8p, i8p, i32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        u8p, f64p]
        L.yo_stub_predict.argtypes = [i8p, ctypes.c_int, f32p, f32p]
        L.yo_stub_key.argtypes = [i8p, ctypes.c_int]
        L.yo_stub_key.restype = ctypes.c_uint64
        L.yo_mcts_search.argtypes = [i8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_float, f64p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                     i32p, f32p, i32p]
        L.yo_mcts_search.restype = ctypes.c_int64
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _boards(boards, n, m):
    b = np.ascontiguousarray(boards, dtype=np.int8).reshape(-1, n * m)
    return b


def legal_mask(boards, players, n, m, rule_flags=0) -> np.ndarray:
    """boards int8[B,n,m] (0/+1/-1), players int8[B] -> uint8[B, n*m] (1 = legal)."""
    b = _boards(boards, n, m)
    p = np.ascontiguousarray(players, dtype=np.int8)
    out = np.zeros((b.shape[0], n * m), dtype=np.uint8)
    lib().yo_legal_mask_batch(_p(b, ctypes.c_int8), _p(p, ctypes.c_int8), b.shape[0], n, m, rule_flags,
                              _p(out, ctypes.c_uint8))
    return out


def next_state(boards, players, actions, n, m, rule_flags=0):
    """Value semantics: returns (new boards int8[B,n,m], next players int8[B])."""
    b = _boards(boards, n, m).copy()
    p = np.ascontiguousarray(players, dtype=np.int8).copy()
    a = np.ascontiguousarray(actions, dtype=np.int32)
    lib().yo_next_state_batch(_p(b, ctypes.c_int8), _p(p, ctypes.c_int8), _p(a, ctypes.c_int32), b.shape[0], n, m,
                              rule_flags)
    return b.reshape(-1, n, m), p


def game_ended(boards, players, n, m, rule_flags=0) -> np.ndarray:
    b = _boards(boards, n, m)
    p = np.ascontiguousarray(players, dtype=np.int8)
    out = np.zeros(b.shape[0], dtype=np.float64)
    lib().yo_game_ended_batch(_p(b, ctypes.c_int8), _p(p, ctypes.c_int8), b.shape[0], n, m, rule_flags,
                              _p(out, ctypes.c_double))
    return out


def env_step(boards, players, actions, n, m, rule_flags=0):
    """mask(side to move) + apply + ended(successor).  Returns (masks, boards', players', results)."""
    b = _boards(boards, n, m).copy()
    p = np.ascontiguousarray(players, dtype=np.int8).copy()
    a = np.ascontiguousarray(actions, dtype=np.int32)
    masks = np.zeros((b.shape[0], n * m), dtype=np.uint8)
    res = np.zeros(b.shape[0], dtype=np.float64)
    lib().yo_env_step_batch(_p(b, ctypes.c_int8), _p(p, ctypes.c_int8), _p(a, ctypes.c_int32), b.shape[0], n, m,
                            rule_flags, _p(masks, ctypes.c_uint8), _p(res, ctypes.c_double))
    return masks, b.reshape(-1, n, m), p, res


def stub_predict(board, n, m):
    """Deterministic hash-stub evaluator: (float32[A] dyadic priors, float32 value)."""
    b = np.ascontiguousarray(board, dtype=np.int8).reshape(-1)
    pol = np.zeros(n * m, dtype=np.float32)
    val = ctypes.c_float(0)
    lib().yo_stub_predict(_p(b, ctypes.c_int8), n * m, _p(pol, ctypes.c_float), ctypes.byref(val))
    return pol, np.float32(val.value)


def mcts_search(board, player, n, m, num_sims, cpuct=1.0, rule_flags=0, noise=None, eps=0.25, evaluator=None):
    """Sequential MCTS (mcts.py:275-343).  evaluator: None = hash stub, else a Python callable
    evaluator(board_int8[n,m]) -> (policy float32[A], value).  Returns dict(counts, child_w, n_evals, stats)."""
    b = np.ascontiguousarray(board, dtype=np.int8).reshape(-1)
    counts = np.zeros(n * m, dtype=np.int32)
    cw = np.zeros(n * m, dtype=np.float32)
    stats = np.zeros(3, dtype=np.int32)
    nz = None
    if noise is not None:
        nz = np.ascontiguousarray(noise, dtype=np.float64)
    cb = None
    if evaluator is not None:
        def _cb(bp, nn, mm, pol, val, ctx):
            arr = np.ctypeslib.as_array(bp, shape=(nn * mm,)).reshape(nn, mm).copy()
            p, v = evaluator(arr)
            np.ctypeslib.as_array(pol, shape=(nn * mm,))[:] = np.asarray(p, dtype=np.float32)
            val[0] = float(np.float32(v))
        cb = EVAL_FN(_cb)
    ne = lib().yo_mcts_search(_p(b, ctypes.c_int8), n, m, int(player), rule_flags, int(num_sims),
                              ctypes.c_float(cpuct), _p(nz, ctypes.c_double) if nz is not None else None,
                              float(eps), ctypes.cast(cb, ctypes.c_void_p) if cb else None, None,
                              _p(counts, ctypes.c_int32), _p(cw, ctypes.c_float), _p(stats, ctypes.c_int32))
    return {"counts": counts, "child_w": cw, "n_evals": int(ne),
            "n_nodes": int(stats[0]), "max_depth": int(stats[1]), "root_visits": int(stats[2])}


# ---------------------------------------------------------------------------------------- dataset / augmentation
def input_planes(boards, n, m) -> np.ndarray:
    """board_to_input (src/yin_yang/ai/neural_network.py:156-196) for int8[N,n,m] boards -> float32[N,5,n,m]."""
    b = _boards(boards, n, m).reshape(-1, n, m)
    occ = b != 0
    x = np.zeros((b.shape[0], 5, n, m), dtype=np.float32)
    x[:, 0] = b == 0
    x[:, 1] = b == 1
    x[:, 2] = b == -1
    x[:, 3] = (occ.sum(axis=2) / m)[:, :, None]        # python float (float64) stored into a float32 tensor
    x[:, 4] = (occ.sum(axis=1) / n)[:, None, :]
    return x


def _symmetries(g: np.ndarray):
    """The 8 forms in the order of DataProcessor.augment_sample (data_utils.py:39-134) of an array [..., n, m]."""
    t = np.swapaxes(g, -1, -2)
    return [g, np.rot90(g, 1, axes=(-2, -1)), np.rot90(g, 2, axes=(-2, -1)), np.rot90(g, 3, axes=(-2, -1)),
            np.flip(g, -1), np.flip(g, -2), t, np.flip(t, (-2, -1))]


def augment_dataset(boards, counts, values, n, m):
    """create_dataset_from_games(..., augment=True) (data_utils.py:182-215) for visit-count policies:
    policy = counts / sum in float64 (uniform 1/A when the sum is 0: mcts.py:209-213), torch.FloatTensor'ed to
    float32 (data_utils.py:34), then for every record the 8 symmetric copies of planes and policy, the value
    replicated.  Returns (planes f32[8N,5,n,m], policies f32[8N,A], values f32[8N]); sample index = 8*record + form."""
    if n != m:
        raise ValueError("the reference's augmentation rotates by 90 degrees: square boards only")
    A = n * m
    c = np.asarray(counts, dtype=np.float64).reshape(-1, A)
    tot = c.sum(axis=1, keepdims=True)
    pi = np.where(tot > 0, c / np.maximum(tot, 1.0), 1.0 / A).astype(np.float32)
    x = input_planes(boards, n, m)
    N = x.shape[0]
    grid = pi.reshape(N, n, m)
    planes = np.stack(_symmetries(x), axis=1).reshape(N * 8, 5, n, m)
    policies = np.stack(_symmetries(grid), axis=1).reshape(N * 8, A)
    vals = np.repeat(np.asarray(values, dtype=np.float64).astype(np.float32), 8)
    return np.ascontiguousarray(planes, dtype=np.float32), np.ascontiguousarray(policies, dtype=np.float32), vals


# ---------------------------------------------------------------------------------------- episode driver
def choice_index(u, p=None, k=None):
    """The index np.random.choice returns for the uniform draw u (numpy/random/mtrand.pyx, legacy `choice`):
    with probabilities p it is cdf.searchsorted(u, side='right') over cdf = cumsum(p) / cumsum(p)[-1]; the recorded
    stream defines the draw without p (np.random.choice(best_moves), self_play.py:146) as floor(u * k)."""
    if p is None:
        return min(int(u * k), k - 1)
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


def play_game(n, m, num_sims, uniforms, noise=None, cpuct=1.0, eps=0.25, temperature_threshold=10, rule_flags=0,
              search_as_black=True, max_plies=None):
    """SelfPlayWorker.play_game (src/yin_yang/ai/self_play.py:72-192) with the hash-stub evaluator and a RECORDED random
    stream: uniforms[step] is the draw behind the action choice of move `step`, `noise` the Dirichlet sample of the
    step-0 root (mcts.py:303-306; None = no noise).  Restated line by line, including the reference's quirks: valid moves
    and the search are always taken for player 1 (:99, :135-137; search_as_black=False searches for the side to move
    instead), the chosen action is applied with the real player (:163, an action illegal for it is silently dropped),
    two consecutive positions without a move for the searched colour end the game (:101-121; 1e-4 if the rules do not
    call it finished), and every example of a game receives the same value (:170-181: the sign is flipped twice per
    example).  Returns (boards int8[E,n,m], pis float64[E,A], zs float64[E], end) with end in {"ended", "passes", "cut"}."""
    A = n * m
    board = np.zeros((n, m), dtype=np.int8)
    player, step, passes = 1, 0, 0
    boards, pis = [], []

    def finish(result, kind):
        zs, value = [], result
        for i in range(len(boards)):                      # self_play.py:112-115 / :172-181
            zs.append(value if i % 2 == 0 else -value)
            value = -value
        return (np.array(boards, dtype=np.int8).reshape(-1, n, m), np.array(pis, dtype=np.float64).reshape(-1, A),
                np.array(zs, dtype=np.float64), kind)

    while True:
        if max_plies is not None and step >= max_plies:
            return finish(float("nan"), "cut")
        temperature = 1.0 if step < temperature_threshold else 0
        sp = 1 if search_as_black else player
        valid = legal_mask(board[None], np.array([sp], np.int8), n, m, rule_flags)[0].astype(np.float64)
        idx = np.flatnonzero(valid == 1)
        if len(idx) == 0:
            passes += 1
            if passes >= 2:
                result = float(game_ended(board[None], np.array([player], np.int8), n, m, rule_flags)[0])
                if result == 0:
                    result = 1e-4
                return finish(result, "passes")
            player = -player
            continue
        passes = 0
        r = mcts_search(board, sp, n, m, num_sims, cpuct=cpuct, rule_flags=rule_flags,
                        noise=noise if (step == 0 and noise is not None) else None, eps=eps)
        counts = r["counts"].astype(np.float64)
        pi = counts / counts.sum() if counts.sum() > 0 else np.ones(A) / A      # mcts.py:207-213 (temperature 1)
        boards.append(board.copy()); pis.append(pi)
        u = float(uniforms[step])
        if temperature == 0:
            best = np.flatnonzero(pi == pi.max())
            action = int(best[choice_index(u, k=len(best))])
        else:
            probs = pi * valid
            if probs.sum() > 0:
                probs = probs / probs.sum()
            else:
                probs = np.zeros_like(valid); probs[idx] = 1.0 / len(idx)
            action = choice_index(u, p=probs)
        nb, npl = next_state(board[None], np.array([player], np.int8), np.array([action], np.int32), n, m, rule_flags)
        board, player = nb[0], int(npl[0])
        step += 1
        result = float(game_ended(board[None], np.array([player], np.int8), n, m, rule_flags)[0])
        if result != 0:
            return finish(result, "ended")
