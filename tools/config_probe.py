"""Developer tool: one self-play move step of an arbitrary configuration (BASELINE.json configs[0] / configs[4] ...).
   YY_N=16 YY_SIMS=1600 YY_GAMES=592 python tools/config_probe.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine
from yinyang_game_alphazero_b200 import network

FLOPS_PER_LEAF = {(8, 8): 380584448, (6, 6): 214014464, (16, 16): 1525481984}
n = m = int(os.environ.get("YY_N", 16))
games = int(os.environ.get("YY_GAMES", 592))
sims = int(os.environ.get("YY_SIMS", 1600))
plies = int(os.environ.get("YY_PLIES", 2))
torch.manual_seed(0)
net = network._Params(n, m, 128, 10).eval()
e = engine.Engine(rows=n, cols=m, n_games=games, n_sims=sims, evaluator="nn", state_dict=net.state_dict(), seed=1)
e.selfplay_run(1)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); e.selfplay_run(plies); ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / plies
st = e.stats()
out = {"board": f"{n}x{m}", "sims": sims, "games": games, "ms_per_move_step": ms, "moves_per_s": games / ms * 1e3,
       "us_per_simulation": ms / (sims + 1) * 1e3, "leaf_evals_per_s": games * (sims + 1) / ms * 1e3,
       "algorithmic_tflops": games * (sims + 1) * FLOPS_PER_LEAF.get((n, m), 0) / ms * 1e3 / 1e12,
       "workspace_gb": e.workspace_bytes / 1e9, "stats": st.__dict__}
print(json.dumps(out))
