"""Developer tool: per-layer timeline (SM clocks) of the tower kernel's CTA 0, first group."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine
from oracle import port

n = m = 8
games = int(os.environ.get("YY_GAMES", 4096))
torch.manual_seed(0)
net = port.build_net(n, m, 128, 10).eval()
e = engine.Engine(rows=n, cols=m, n_games=games, n_sims=4, evaluator="nn", state_dict=net.state_dict())
black, white, players = engine.random_playout(games, torch.arange(games, dtype=torch.int32) % 50, n, m)
for _ in range(int(os.environ.get("YY_WARM", 1))):
    e.evaluate(black, white)
dbg = torch.zeros(1024, dtype=torch.int64, device="cuda")
e.L.yy_engine_set_debug_stamps(e.handle, ctypes.c_void_p(dbg.data_ptr()))
e.evaluate(black, white)
torch.cuda.synchronize()
e.L.yy_engine_set_debug_stamps(e.handle, None)
raw = dbg.cpu().numpy()
d = raw[:96].reshape(-1, 4)
t0 = d[0, 0]
print("layer  mma_issue_start  mma_issue_len  acc_ready(epi start)  epi_len   layer_total(from issue start to epi end)")
for l in range(22):
    print(f"{l:3d} {d[l,0]-t0:10d} {d[l,1]-d[l,0]:10d} {d[l,2]-t0:10d} {d[l,3]-d[l,2]:10d} {d[l,3]-d[l,0]:10d}   mma_phase(issue start->acc ready)={d[l,2]-d[l,0]}")

import numpy as np
gt = raw[128:128 + 2 * 148].reshape(-1, 2)
busy = gt[:, 0] > 0
t00 = gt[busy, 0].min()
print("CTA start (us after first):", np.percentile((gt[busy, 0] - t00) / 1e3, [0, 50, 100]))
print("CTA end   (us after first start):", np.percentile((gt[busy, 1] - t00) / 1e3, [0, 5, 50, 95, 100]))
print("CTA duration us:", np.percentile((gt[busy, 1] - gt[busy, 0]) / 1e3, [0, 50, 100]), "busy CTAs", busy.sum())
for name, off in (("cta0", 512), ("cta100", 528)):
    g = raw[off:off + 8]
    g = g[g > 0]
    print(name, "group start clocks (delta):", np.diff(g), "first group start", g[0] if len(g) else None)
print("cta0 kernel-end clock - first group start:", raw[544] - raw[512], " cta100:", raw[544 + 100] - raw[528])
