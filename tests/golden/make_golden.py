#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED Python reference.

Run in the build container only (it imports /root/reference, which does not exist on
the GPU box):

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden.py

What is pinned
  rules_<n>x<m>.npz  boards (random legal play AND arbitrary random fills), both colours:
                     YinYangGame.getValidMoves, getNextState (legal, illegal and occupied
                     actions), getGameEnded   (yin_yang_game.py:39-110, yin_yang_logic.py)
  mcts_*.npz         root.get_children_visit_counts() and child value sums of the unmodified
                     src/yin_yang/ai/mcts.py driven through a value-semantics Game adapter
                     (getNextState deep-copies) with the deterministic hash-stub
                     evaluator (dyadic priors / values)
  augment_*.npz      data_utils.create_dataset_from_games (preprocess_sample + augment_sample: the 8 symmetric
                     copies of the input planes and of the policy, and the replicated value) on random-play
                     boards with visit-count policies   (data_utils.py:16-215)
  net_*.npz          a small YinYangNeuralNetwork (random weights AND random BN statistics):
                     state_dict, board_to_input planes, forward logits/value, predict()
The reference modules open *.log files in the CWD on import -> run from a scratch dir.
"""
import copy
import logging
import os
import sys

import numpy as np

REF = os.environ.get("YY_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
logging.disable(logging.CRITICAL)

from src.yin_yang import YinYangGame  # noqa: E402
from src.yin_yang.ai.mcts import MCTS  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
M64 = (1 << 64) - 1


# ------------------------------------------------------------------ hash stub (pure Python ints)
def mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def stub_key(arr):
    flat = np.asarray(arr).reshape(-1)
    cells = flat.size
    words = (cells + 63) // 64
    key = 0x243F6A8885A308D3
    for want in (1, -1):
        for w in range(words):
            bits = 0
            for k in range(64):
                i = w * 64 + k
                if i < cells and flat[i] == want:
                    bits |= 1 << k
            key = mix64(key ^ bits)
    return key


class HashStub:
    """predict(board) -> (np.float32[A], np.float32): same contract as neural_network.py:125-154."""

    def __init__(self):
        self.calls = 0

    def predict(self, board):
        self.calls += 1
        arr = board.get_board()
        key = stub_key(arr)
        A = arr.size
        pol = np.empty(A, dtype=np.float32)
        for a in range(A):
            pol[a] = np.float32(1 + (mix64((key + 0xD1B54A32D192ED03 * (a + 1)) & M64) >> 52)) / np.float32(65536.0)
        val = np.float32(int(mix64(key ^ 0xA0761D6478BD642F) >> 47) - 65536) / np.float32(65536.0)
        return pol, val


class CopyGame(YinYangGame):
    """Value-semantics adapter through the reference's own duck-typed Game seam (mcts.py:231,249)."""

    def getNextState(self, board, player, action):
        return super().getNextState(copy.deepcopy(board), player, action)


# ------------------------------------------------------------------ rules fixtures
def random_play_board(game, rng, plies):
    b = game.getInitBoard()
    player = 1
    for _ in range(plies):
        mask = game.getValidMoves(b, player)
        idx = np.flatnonzero(mask)
        if idx.size == 0:
            player = -player
            continue
        b, player = game.getNextState(b, player, int(rng.choice(idx)))
    return b, player


def arbitrary_board(game, rng, fill):
    b = game.getInitBoard()
    r = rng.random((game.n, game.m))
    b.board[:] = np.where(r < fill / 2, 1, np.where(r < fill, -1, 0)).astype(np.int8)
    return b


def make_rules(n, m, n_play, n_arb, seed):
    game = YinYangGame(n, m)
    rng = np.random.default_rng(seed)
    A = n * m
    boards, players, mask_b, mask_w, actions, next_boards, next_players, ended_b, ended_w = ([] for _ in range(9))
    specials = []
    e = game.getInitBoard(); specials.append((e, 1))                                  # empty board
    s = game.getInitBoard(); s.board[0, 0] = 1; s.board[n - 1, m - 1] = 1; specials.append((s, 1))  # 2 components
    q = game.getInitBoard(); q.board[0:2, 0:2] = -1; specials.append((q, -1))         # pre-existing 2x2
    f = game.getInitBoard(); f.board[:] = 1; f.board[::2, ::2] = -1; specials.append((f, 1))  # full board
    max_plies = int(A * 0.95)
    for i in range(n_play):
        specials.append(random_play_board(game, rng, int(rng.integers(0, max_plies + 1))))
    for i in range(n_arb):
        specials.append((arbitrary_board(game, rng, float(rng.uniform(0.05, 0.95))), int(rng.choice([1, -1]))))
    for b, player in specials:
        arr = b.get_board()
        mb = game.getValidMoves(b, 1)
        mw = game.getValidMoves(b, -1)
        mine = mb if player == 1 else mw
        legal = np.flatnonzero(mine)
        u = rng.random()
        if legal.size and u < 0.6:
            a = int(rng.choice(legal))
        else:
            a = int(rng.integers(0, A))          # often illegal / occupied -> silent no-op
        nb, npl = game.getNextState(copy.deepcopy(b), player, a)
        boards.append(arr); players.append(player)
        mask_b.append(mb.astype(np.uint8)); mask_w.append(mw.astype(np.uint8))
        actions.append(a); next_boards.append(nb.get_board()); next_players.append(npl)
        ended_b.append(float(game.getGameEnded(b, 1))); ended_w.append(float(game.getGameEnded(b, -1)))
    np.savez_compressed(
        os.path.join(OUT, f"rules_{n}x{m}.npz"),
        n=n, m=m, boards=np.array(boards, dtype=np.int8), players=np.array(players, dtype=np.int8),
        mask_black=np.array(mask_b), mask_white=np.array(mask_w), actions=np.array(actions, dtype=np.int32),
        next_boards=np.array(next_boards, dtype=np.int8), next_players=np.array(next_players, dtype=np.int8),
        ended_black=np.array(ended_b), ended_white=np.array(ended_w))
    print(f"rules {n}x{m}: {len(boards)} boards")


# ------------------------------------------------------------------ MCTS fixtures
def make_mcts(name, n, m, sims, plies=0, seed=0, noise=False, cpuct=1.0):
    game = CopyGame(n, m)
    rng = np.random.default_rng(seed)
    board, player = random_play_board(game, rng, plies)
    board = copy.deepcopy(board)
    start = board.get_board()
    stub = HashStub()
    mcts = MCTS(game, stub, num_simulations=sims, cpuct=cpuct, dirichlet_noise=noise, verbose=0)
    noise_used = np.zeros(0)
    if noise:
        k = int(game.getValidMoves(board, player).sum())
        noise_used = np.random.default_rng(seed + 99).dirichlet([0.3] * k)
        orig = np.random.dirichlet
        np.random.dirichlet = lambda alpha, size=None: noise_used.copy()
    try:
        probs, root = mcts.search(board, player, add_exploration_noise=noise)
    finally:
        if noise:
            np.random.dirichlet = orig
    assert np.array_equal(board.get_board(), start), "search mutated the root board"
    counts = root.get_children_visit_counts().astype(np.int32)
    child_w = np.zeros(n * m, dtype=np.float32)
    wdtype = set()
    for a, ch in root.children.items():
        child_w[a] = np.float32(ch.value_sum)
        wdtype.add(type(ch.value_sum).__name__)
    np.savez_compressed(os.path.join(OUT, f"mcts_{name}.npz"), n=n, m=m, sims=sims, player=player, cpuct=cpuct,
                        board=start, counts=counts, child_w=child_w, probs=probs, noise=noise_used,
                        n_evals=stub.calls, root_visits=root.visits)
    print(f"mcts {name}: evals={stub.calls} root_visits={root.visits} top={np.sort(counts)[-3:]} Wtypes={wdtype}")


# ------------------------------------------------------------------ network fixtures
def make_net(name, n, m, channels, blocks, seed):
    import torch
    from src.yin_yang.ai.neural_network import YinYangNeuralNetwork
    game = YinYangGame(n, m)
    torch.manual_seed(seed)
    net = YinYangNeuralNetwork(game, num_channels=channels, num_res_blocks=blocks)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):   # non-trivial BN so that folding is really tested
                mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.2)
                mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 1.5 + 0.25)
                mod.weight.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            if isinstance(mod, (torch.nn.Conv2d, torch.nn.Linear)):
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.05)
    net.eval()
    rng = np.random.default_rng(seed)
    boards, planes, logits, values, policies, pvalues = [], [], [], [], [], []
    for i in range(24):
        b, _ = random_play_board(game, rng, int(rng.integers(0, n * m)))
        x = net.board_to_input(b)
        with torch.no_grad():
            lg, v = net.forward(x.unsqueeze(0))
        p, pv = net.predict(b)
        boards.append(b.get_board()); planes.append(x.numpy()); logits.append(lg.numpy()[0]); values.append(v.numpy()[0, 0])
        policies.append(p); pvalues.append(pv)
    sd = {k: v.numpy() for k, v in net.state_dict().items()}
    np.savez_compressed(os.path.join(OUT, f"net_{name}.npz"), n=n, m=m, channels=channels, blocks=blocks,
                        boards=np.array(boards, dtype=np.int8), planes=np.array(planes, dtype=np.float32),
                        logits=np.array(logits, dtype=np.float32), values=np.array(values, dtype=np.float32),
                        policies=np.array(policies, dtype=np.float32), pvalues=np.array(pvalues, dtype=np.float32),
                        **{"sd." + k: v for k, v in sd.items()})
    print(f"net {name}: params={sum(v.size for v in sd.values())}")


def make_augment(name, n, m, count, seed):
    from src.yin_yang.ai.data_utils import create_dataset_from_games
    rng = np.random.default_rng(seed)
    game = YinYangGame(n, m)
    boards, counts, values, data = [], [], [], []
    for i in range(count):
        b, player = random_play_board(game, rng, int(rng.integers(0, n * m)))
        mask = game.getValidMoves(b, player)
        c = (rng.integers(0, 200, size=n * m) * (rng.random(n * m) < 0.6) * (mask > 0)).astype(np.uint16)
        if i % 7 == 3:
            c[:] = 0                                              # no visits at all: uniform fallback (mcts.py:211-213)
        tot = float(c.sum())
        pi = c.astype(np.float64) / tot if tot > 0 else np.ones(n * m) / (n * m)   # get_children_distribution, temperature 1
        z = float(rng.choice([1.0, -1.0, 0.0001, -0.0001]))
        boards.append(b.get_board().copy()); counts.append(c); values.append(z)
        data.append((b, pi, z))
    bt, pt, vt = create_dataset_from_games(data, game, augment=True)
    np.savez_compressed(os.path.join(OUT, f"augment_{name}.npz"), n=n, m=m, boards=np.array(boards, dtype=np.int8),
                        counts=np.array(counts, dtype=np.uint16), values=np.array(values, dtype=np.float64),
                        planes=np.stack([t.numpy() for t in bt]).astype(np.float32),
                        policies=np.stack([t.numpy() for t in pt]).astype(np.float32),
                        out_values=np.stack([t.numpy() for t in vt]).astype(np.float32)[:, 0])
    print(f"augment {name}: {count} records -> {len(bt)} samples")


if __name__ == "__main__":
    which = sys.argv[1:] or ["rules", "mcts", "net", "augment"]
    if "augment" in which:
        make_augment("6x6", 6, 6, 24, 21)
        make_augment("8x8", 8, 8, 40, 22)
        make_augment("16x16", 16, 16, 6, 23)
    if "rules" in which:
        make_rules(4, 4, 120, 120, 1)
        make_rules(6, 6, 200, 150, 2)
        make_rules(8, 8, 300, 200, 3)
        make_rules(5, 7, 80, 80, 4)
        make_rules(16, 16, 24, 24, 5)
    if "mcts" in which:
        make_mcts("4x4_s300", 4, 4, 300)
        make_mcts("4x4_s300_mid", 4, 4, 300, plies=6, seed=7)
        make_mcts("6x6_s400", 6, 6, 400)
        make_mcts("6x6_s100_noise", 6, 6, 100, noise=True, seed=3)
        make_mcts("8x8_s800", 8, 8, 800)
        make_mcts("8x8_s200_mid_white", 8, 8, 200, plies=21, seed=11)
        make_mcts("8x8_s200_late", 8, 8, 200, plies=44, seed=5, cpuct=1.5)
        make_mcts("16x16_s120", 16, 16, 120)
    if "net" in which:
        make_net("6x6_c16_b2", 6, 6, 16, 2, 0)
        make_net("4x4_c16_b1", 4, 4, 16, 1, 1)
