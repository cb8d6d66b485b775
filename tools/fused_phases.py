"""Developer tool: where a persistent-search CTA spends its cycles (tower / FC heads / softmax+tree / barrier / zero)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine
from yinyang_game_alphazero_b200 import network

n = m = 8
games = int(os.environ.get("YY_GAMES", 4096))
sims = int(os.environ.get("YY_SIMS", 200))
torch.manual_seed(0)
net = network._Params(n, m, 128, 10).eval()
e = engine.Engine(rows=n, cols=m, n_games=games, n_sims=sims, evaluator="nn", state_dict=net.state_dict(), seed=1)
e.selfplay_run(int(os.environ.get("YY_PLIES", 3)))
torch.cuda.synchronize()
dbg = torch.zeros(1024, dtype=torch.int64, device="cuda")
e.L.yy_engine_set_debug_stamps(e.handle, ctypes.c_void_p(dbg.data_ptr()))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); e.selfplay_run(1); ev1.record()
torch.cuda.synchronize()
e.L.yy_engine_set_debug_stamps(e.handle, None)
raw = dbg.cpu().numpy()
ms = ev0.elapsed_time(ev1)
print(f"move step {ms:.2f} ms = {ms / (sims + 1) * 1e3:.1f} us per simulation; stats {e.stats()}")
gt = raw[128:128 + 2 * 148].reshape(-1, 2)
busy = gt[:, 0] > 0
print("CTA duration ms:", np.percentile((gt[busy, 1] - gt[busy, 0]) / 1e6, [0, 50, 100]), "busy CTAs", busy.sum())
for name, off in (("cta0", 600), ("cta100", 760)):
    ph = raw[off:off + 128].reshape(16, 8)
    tot = ph[:, :5].sum(axis=1)
    print(name, "per-iteration cycles by phase (warp 0 / warp 15): tower, fc, heads+tree, barrier, zero | inside fc: panel loads, MMA waits, scatter")
    for w in (0, 15):
        print("   ", (ph[w] / (sims + 1)).round(0), "total", round(tot[w] / (sims + 1)))
