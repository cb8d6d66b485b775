"""Developer tool: where the time of one yy_env_step launch goes -- launches of 64 / 4,096 / 65,536 / 1 M boards on empty boards
(no flood fill at all: the launch + memory floor) and on random-play boards, 8 launches per CUDA-graph replay, L2 flushed."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine, bitboard

R = C = int(os.environ.get("YY_SIDE", 8))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = {}
for n in (64, 4096, 65536, 1 << 20):
    for kind in ("empty", "random_play"):
        plies = (torch.arange(n, dtype=torch.int32) % (R * C - 12)) if kind == "random_play" else torch.zeros(n, dtype=torch.int32)
        b0, w0, p0 = engine.random_playout(n, plies, R, C, seed=0xC0FFEE)
        mask0 = engine.legal_mask(b0, w0, p0, R, C)
        bits = bitboard.unpack_bits(mask0.cpu().numpy().view(np.uint64), R, C)
        acts = torch.from_numpy(np.where(bits.any(axis=1), bits.argmax(axis=1), -1).astype(np.int32)).cuda()
        om, orr = torch.empty_like(b0), torch.empty_like(p0)
        bufs = [(b0.clone(), w0.clone(), p0.clone()) for _ in range(8)]
        engine.env_step(*bufs[0], acts, R, C, out_mask=om, out_result=orr)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for b, w, p in bufs:
                engine.env_step(b, w, p, acts, R, C, out_mask=om, out_result=orr)
        ms = []
        for r in range(12):
            for b, w, p in bufs:
                b.copy_(b0); w.copy_(w0); p.copy_(p0)
            flush.fill_(r); torch.cuda.synchronize()
            ev0.record(); g.replay(); ev1.record(); torch.cuda.synchronize()
            ms.append(ev0.elapsed_time(ev1))
        us = float(np.median(ms)) / 8 * 1e3
        out[f"{n}_{kind}"] = {"us_per_launch": round(us, 3), "G_steps_per_s": round(n / us / 1e3, 3)}
print(json.dumps(out))
