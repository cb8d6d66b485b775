"""Developer tool: one small invocation of every kernel family at odd sizes (partial blocks, non-square and 16x16
boards, remainder batches) -- a quick crash / overflow check after a kernel change, and the driver for
`compute-sanitizer --tool memcheck python tools/sanitize_smoke.py [part ...]` where the sanitizer may be run (it is closed
on the round-1 GPU pool: gpurun refuses it).  Correctness of the results is the job of tests/.
Parts: rules, dataset, tree, nn, selfplay, learner (default: all)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yy_b200  # noqa: F401,E402
from yinyang_game_alphazero_b200 import engine  # noqa: E402


def boards_after_random_play(n, m, count, seed):
    plies = torch.arange(count, dtype=torch.int32, device="cuda") % (n * m - n)
    b, w, p = engine.random_playout(count, plies, n, m, seed=seed)
    nb, _ = engine.unpack_boards_dev(n, m, b, w)
    return nb, p.cpu().numpy()


def part_rules():
    for (n, m, flags) in [(8, 8, 0), (16, 16, 0), (5, 7, 1), (6, 6, 0)]:
        boards, players = boards_after_random_play(n, m, 1031, seed=3)          # odd count: partial blocks
        acts = np.random.default_rng(0).integers(-1, n * m + 1, size=len(players)).astype(np.int32)
        engine.legal_mask_host(boards, players, n, m, flags)
        engine.env_step_host(boards, players, acts, n, m, flags)
        engine.next_state_host(boards, players, acts, n, m, flags)
        engine.ended_host(boards, players, n, m, flags)
    print("rules ok")


def part_dataset():
    rng = np.random.default_rng(1)
    for n, count in [(8, 517), (6, 33), (16, 9)]:
        boards = rng.integers(-1, 2, size=(count, n, n)).astype(np.int8)
        counts = rng.integers(0, 800, size=(count, n * n)).astype(np.uint16)
        engine.augment_samples_host(boards, n, n, counts=counts, values=rng.choice([1.0, -1.0, 0.0001], size=count))
    print("dataset ok")


def part_tree():
    for (n, games, sims, kw) in [(8, 24, 64, {}), (8, 24, 64, {"step_kernels": True}), (6, 8, 48, {"leaves_per_step": 4}),
                                 (16, 4, 24, {})]:
        boards, players = boards_after_random_play(n, n, games, seed=5)
        e = engine.Engine(rows=n, cols=n, n_games=games, n_sims=sims, evaluator="stub", **kw)
        counts, _ = e.search_host(boards, players)
        assert e.stats().overflow == 0 and counts.sum() > 0
        e.close()
    print("tree ok")


def _net(n, ch, blocks):
    from yinyang_game_alphazero_b200 import network
    torch.manual_seed(0)
    return network._Params(n, n, ch, blocks).eval()   # reference layout + initialisation; only its state_dict is used


def part_nn():
    for (n, games, sims, kw) in [(8, 12, 12, {}), (8, 12, 8, {"step_kernels": True}), (6, 8, 10, {}), (8, 2, 10, {"leaves_per_step": 4})]:
        net = _net(n, 128, 2)
        boards, players = boards_after_random_play(n, n, games, seed=7)
        e = engine.Engine(rows=n, cols=n, n_games=games, n_sims=sims, evaluator="nn", state_dict=net.state_dict(), **kw)
        e.evaluate_host(boards, want_logits=True)
        counts, _ = e.search_host(boards, players)
        assert e.stats().overflow == 0 and counts.sum() > 0
        e.close()
    print("nn ok")


def part_selfplay():
    net = _net(6, 128, 1)
    e = engine.Engine(rows=6, cols=6, n_games=8, n_sims=12, evaluator="nn", state_dict=net.state_dict())
    e.selfplay_run(40)                            # long enough for games to end and slots to restart
    st = e.stats()
    assert st.moves == 8 * 40 and st.overflow == 0
    e.replay()
    e.close()
    print("selfplay ok", st)


def part_learner():
    from yinyang_game_alphazero_b200 import learner
    for (n, ch, blocks, batch) in [(8, 128, 1, 8), (6, 32, 2, 5)]:
        net = _net(n, ch, blocks)
        L = learner.Learner(n, n, ch, blocks, batch_size=batch, state_dict=net.state_dict())
        planes = torch.rand(batch, 5, n, n, device="cuda")
        pi = torch.softmax(torch.randn(batch, n * n, device="cuda"), 1)
        z = torch.rand(batch, device="cuda") * 2 - 1
        for _ in range(3):                        # eager step, graph capture, graph replay
            losses = L.step(planes, pi, z)
        assert torch.isfinite(losses).all()
    print("learner ok")


if __name__ == "__main__":
    parts = sys.argv[1:] or ["rules", "dataset", "tree", "nn", "selfplay", "learner"]
    torch.cuda.set_device(0)
    for p in parts:
        globals()["part_" + p]()
    torch.cuda.synchronize()
    print("sanitize_smoke done:", " ".join(parts))
