#!/usr/bin/env python
"""Golden fixtures of the EPISODE DRIVER, from the unmodified reference: SelfPlayWorker.play_game
(src/yin_yang/ai/self_play.py:72-192) is run as it stands, with
  * the value-semantics Game adapter and the hash-stub evaluator of make_golden.py (the MCTS itself is the
    unmodified src/yin_yang/ai/mcts.py), and
  * np.random.choice / np.random.dirichlet -- the only two draws play_game and MCTS.search make -- replaced by a
    RECORDED stream: one uniform per move (choice_index below = numpy's own cdf.searchsorted(u, side='right') for the
    draw with probabilities; floor(u*k) for np.random.choice(best_moves)) and one Dirichlet sample for the step-0 root.
Stored per game: every example (board, pi float64, z), the uniforms, the noise vector and how the game ended.

Run in the build container only:   cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden_selfplay.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import CopyGame, HashStub, OUT  # noqa: E402  (puts /root/reference on sys.path)

from src.yin_yang.ai.mcts import MCTS  # noqa: E402
from src.yin_yang.ai.self_play import SelfPlayWorker  # noqa: E402


def choice_index(u, p=None, k=None):
    if p is None:
        return min(int(u * k), k - 1)
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


class RecordedRandom:
    """Stands in for np.random.choice / np.random.dirichlet while one game is played."""

    def __init__(self, seed, alpha):
        self.rng = np.random.default_rng(seed)
        self.uniforms, self.noise, self.alpha = [], None, alpha

    def choice(self, a, size=None, replace=True, p=None):
        assert size is None
        u = float(self.rng.random())
        self.uniforms.append(u)
        if p is None:
            a = np.asarray(a)
            return a[choice_index(u, k=len(a))]
        assert isinstance(a, (int, np.integer))
        return choice_index(u, p=p)

    def dirichlet(self, alpha, size=None):
        assert self.noise is None and size is None, "one Dirichlet draw per game (step 0 only)"
        assert all(x == self.alpha for x in alpha)
        self.noise = self.rng.dirichlet(alpha)
        return self.noise.copy()


def make_game(name, n, m, sims, seed, temperature_threshold=10, cpuct=1.0, alpha=0.3, eps=0.25):
    game = CopyGame(n, m)
    w = SelfPlayWorker.__new__(SelfPlayWorker)            # __init__ only loads a network file and builds the MCTS
    w.game, w.num_simulations, w.num_games = game, sims, 1
    w.temperature_threshold = temperature_threshold
    w.neural_net = HashStub()
    w.mcts = MCTS(game=game, neural_net=w.neural_net, num_simulations=sims, cpuct=cpuct, dirichlet_alpha=alpha,
                  dirichlet_epsilon=eps, num_threads=1)
    rec = RecordedRandom(seed, alpha)
    orig = np.random.choice, np.random.dirichlet
    np.random.choice, np.random.dirichlet = rec.choice, rec.dirichlet
    try:
        examples = w.play_game()
    finally:
        np.random.choice, np.random.dirichlet = orig
    boards = np.array([np.asarray(ex[0].get_board()) for ex in examples], dtype=np.int8)
    pis = np.array([ex[1] for ex in examples], dtype=np.float64)
    zs = np.array([ex[2] for ex in examples], dtype=np.float64)
    assert len(rec.uniforms) == len(examples)
    stones = np.abs(boards).sum(axis=(1, 2))
    dropped = int((np.diff(stones) == 0).sum())           # moves the real player could not make (silently dropped)
    np.savez_compressed(os.path.join(OUT, f"selfplay_{name}.npz"), n=n, m=m, sims=sims, cpuct=cpuct, alpha=alpha, eps=eps,
                        temperature_threshold=temperature_threshold, boards=boards, pis=pis, zs=zs,
                        uniforms=np.array(rec.uniforms), noise=rec.noise if rec.noise is not None else np.zeros(0))
    print(f"selfplay {name}: {len(examples)} examples, z={sorted(set(zs.tolist()))}, dropped moves={dropped}, "
          f"final stones={int(stones[-1])}")
    return zs


if __name__ == "__main__":
    # chosen (by running a seed sweep) to cover: win / loss / the 1e-4 "draw" of the double-pass branch, games past the
    # temperature threshold (argmax branch with ties), silently dropped moves, a 6x6 game well past ply 10, an 8x8 game
    make_game("4x4_a", 4, 4, 60, seed=1, temperature_threshold=3)
    make_game("4x4_b", 4, 4, 60, seed=2, temperature_threshold=3)
    make_game("4x4_c", 4, 4, 40, seed=5, temperature_threshold=10, cpuct=1.5)
    make_game("5x7", 5, 7, 40, seed=3, temperature_threshold=4)
    make_game("6x6_a", 6, 6, 50, seed=4)
    make_game("6x6_b", 6, 6, 80, seed=7, temperature_threshold=6, eps=0.0)
    make_game("8x8", 8, 8, 60, seed=6)
    # games the rules end after a move (self_play.py:167-188): result -1 / +1 for every example
    make_game("4x4_loss", 4, 4, 30, seed=201)
    make_game("4x4_win", 4, 4, 30, seed=235)
    make_game("5x5_win", 5, 5, 40, seed=242, temperature_threshold=4)
    make_game("6x6_loss", 6, 6, 40, seed=200, temperature_threshold=6)
