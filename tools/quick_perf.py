"""Quick device-side timing probe (developer tool, not the bench): NN forward and a short self-play burst."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine
from yinyang_game_alphazero_b200 import network

FLOPS_PER_LEAF = {(8, 8): 380584448, (6, 6): 214014464, (16, 16): 1525481984}


def main():
    n = m = int(os.environ.get("YY_N", 8))
    games = int(os.environ.get("YY_GAMES", 4096))
    sims = int(os.environ.get("YY_SIMS", 64))
    torch.manual_seed(0)
    net = network._Params(n, m, 128, 10).eval()
    e = engine.Engine(rows=n, cols=m, n_games=games, n_sims=sims, evaluator="nn", state_dict=net.state_dict(), seed=1)
    plies = torch.arange(games, dtype=torch.int32) % (n * m * 4 // 5)
    black, white, players = engine.random_playout(games, plies, n, m)
    out = {}
    for _ in range(3):
        e.evaluate(black, white)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    ev0.record()
    for _ in range(reps):
        e.evaluate(black, white)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    out["nn_forward_ms"] = ms
    out["nn_boards_per_s"] = games / ms * 1e3
    out["nn_tflops"] = games * FLOPS_PER_LEAF.get((n, m), 0) / ms * 1e3 / 1e12
    # short self-play burst
    e.selfplay_run(1); torch.cuda.synchronize()
    ev0.record(); e.selfplay_run(2); ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    st = e.stats()
    out["selfplay_ms_per_move_step"] = ms / 2
    out["selfplay_moves_per_s_at_sims"] = games * 2 / ms * 1e3
    out["sims"] = sims
    out["stats"] = st.__dict__
    # stub-mode tree-only throughput
    e2 = engine.Engine(rows=n, cols=m, n_games=games, n_sims=sims, evaluator="stub", seed=1)
    e2.selfplay_run(1); torch.cuda.synchronize()
    ev0.record(); e2.selfplay_run(2); ev1.record(); torch.cuda.synchronize()
    out["tree_only_ms_per_sim_step"] = ev0.elapsed_time(ev1) / 2 / (sims + 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
