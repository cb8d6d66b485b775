"""Developer tool: latency of a lock-step search (400 simulations, 128x10 network) for small numbers of positions -- the arena's and the
players' case; shows the effect of spreading a small batch over the CTAs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import yy_b200
from yinyang_game_alphazero_b200 import engine, network
torch.manual_seed(0)
sd = network._Params(8, 8, 128, 10).state_dict()
for games in (16, 64, 300, 700, 1184):
    e = engine.Engine(rows=8, cols=8, n_games=games, n_sims=400, evaluator="nn", state_dict=sd, seed=1)
    boards = np.zeros((games, 8, 8), np.int8); players = np.ones(games, np.int8)
    e.search_host(boards, players); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): e.search_host(boards, players)
    print(games, "games, 400 sims: ms per lock-step search", round((time.perf_counter() - t0) / 3 * 1e3, 1))
    e.close()
