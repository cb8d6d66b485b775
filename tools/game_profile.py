"""Developer tool: ms per self-play move step as the games progress (8x8; all games start together)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine
from yinyang_game_alphazero_b200 import network

n = m = 8
games = int(os.environ.get("YY_GAMES", 4096)); sims = int(os.environ.get("YY_SIMS", 200)); plies = int(os.environ.get("YY_PLIES", 70))
torch.manual_seed(0)
net = network._Params(n, m, 128, 10).eval()
e = engine.Engine(rows=n, cols=m, n_games=games, n_sims=sims, evaluator="nn", state_dict=net.state_dict(), seed=1,
                  replay_capacity=games * (plies + 2))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(plies + 1)]
ev[0].record()
stats = []
for p in range(plies):
    e.selfplay_run(1)
    ev[p + 1].record()
torch.cuda.synchronize()
st = e.stats()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(plies)]
print("ms per move step by ply:", [round(x, 1) for x in ms])
print("total", round(sum(ms), 1), "ms; evals per sim-slot:", st.evals / max(1, st.sims + st.moves), st)
