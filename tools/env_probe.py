"""Developer tool: BASELINE.json configs[1] (65,536 random-play 8x8 boards) -- time yy_env_step launches back to back."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine, bitboard

N = int(os.environ.get("YY_BOARDS", 65536)); R = C = 8
plies = torch.arange(N, dtype=torch.int32) % 52
b0, w0, p0 = engine.random_playout(N, plies, R, C, seed=0xC0FFEE)
mask0 = engine.legal_mask(b0, w0, p0, R, C)
bits = bitboard.unpack_bits(mask0.cpu().numpy().view(np.uint64), R, C)
acts = torch.from_numpy(np.where(bits.any(axis=1), bits.argmax(axis=1), -1).astype(np.int32)).cuda()
om, orr = torch.empty_like(b0), torch.empty_like(p0)
reps = int(os.environ.get("YY_REPS", 50))
bufs = [(b0.clone(), w0.clone(), p0.clone()) for _ in range(reps)]
engine.env_step(*bufs[0], acts, R, C, out_mask=om, out_result=orr)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for b, w, p in bufs[1:]:
    engine.env_step(b, w, p, acts, R, C, out_mask=om, out_result=orr)
ev1.record(); torch.cuda.synchronize()
us = ev0.elapsed_time(ev1) / (reps - 1) * 1e3
print(json.dumps({"boards": N, "us_per_launch": us, "steps_per_s": N / us * 1e6, "GBps_algorithmic": 52 * N / us / 1e3}))
