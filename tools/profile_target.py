"""ncu target: rolling self-play of BASELINE.json configs[2] (8x8, 800 sims, 4,096 games, 128x10 network): three warm-up
launches of 801 evaluation steps, then ONE launch of YY_PROFILE_ITERS (default 120) steps -- the launch to capture with
    ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 3 -c 1 -o gpurun_out/prof python tools/profile_target.py
(mid-game trees, every slot busy; a full 801-step launch profiles the same code for 6x longer)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine, network

n = int(os.environ.get("YY_N", 8))
sims = int(os.environ.get("YY_SIMS", 800))
games = int(os.environ.get("YY_GAMES", 4096))
evaluator = os.environ.get("YY_EVAL", "nn")
torch.manual_seed(0)
sd = network._Params(n, n, 128, 10).state_dict() if evaluator == "nn" else None
e = engine.Engine(rows=n, cols=n, n_games=games, n_sims=sims, evaluator=evaluator, state_dict=sd, seed=1, replay_capacity=games * 32)
for _ in range(3):
    e.selfplay_advance(sims + 1)
torch.cuda.synchronize()
s0 = e.stats()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = int(os.environ.get("YY_PROFILE_ITERS", 120))
ev0.record(); e.selfplay_advance(iters); ev1.record(); torch.cuda.synchronize()
s1 = e.stats()
print(f"profiled launch: {iters} steps, {ev0.elapsed_time(ev1):.2f} ms, {s1.tower_evals - s0.tower_evals} evaluations, {s1.moves - s0.moves} moves, "
      f"{s1.sims - s0.sims} simulations")
e.close()
