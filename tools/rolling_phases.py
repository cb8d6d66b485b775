"""Developer tool: rolling self-play (yy_selfplay_advance) -- moves/s, evaluations per move and where a CTA of the
persistent kernel spends its cycles, for several values of descents_per_step after a given number of warm-up steps.

    python tools/rolling_phases.py [descents ...]        env: YY_GAMES, YY_SIMS, YY_WARM, YY_STEPS, YY_N
"""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine, network

n = m = int(os.environ.get("YY_N", 8))
games = int(os.environ.get("YY_GAMES", 4096))
sims = int(os.environ.get("YY_SIMS", 800))
warm = int(os.environ.get("YY_WARM", 5))
steps = int(os.environ.get("YY_STEPS", 4))
FLOPS = {8: 380584448, 6: 214014464, 16: 1525481984}[n]
torch.manual_seed(0)
sd = network._Params(n, m, 128, 10).state_dict()
for desc in [int(x) for x in sys.argv[1:]] or [0]:
    e = engine.Engine(rows=n, cols=m, n_games=games, n_sims=sims, evaluator="nn", state_dict=sd, seed=1, descents_per_step=desc,
                      replay_capacity=games * 3 * (warm + steps + 2))
    iters = sims + 1
    e.L.yy_engine_set_debug_flags(e.handle, int(os.environ.get("YY_DBG_FLAGS", 0)))
    for _ in range(warm):
        e.selfplay_advance(iters)
    torch.cuda.synchronize()
    s0 = e.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        e.selfplay_advance(iters)
    ev1.record(); torch.cuda.synchronize()
    s1 = e.stats()
    ms = ev0.elapsed_time(ev1)
    dm, de = s1.moves - s0.moves, s1.tower_evals - s0.tower_evals
    out = {"descents_per_step": desc, "ms_per_step": ms / steps, "us_per_iteration": ms / steps / iters * 1e3, "moves_per_s": dm / ms * 1e3,
           "evals_per_move": de / max(1, dm), "slot_occupancy": de / (steps * iters * games), "tflops": de * FLOPS / ms / 1e9,
           "games_finished": s1.games_finished, "max_depth": s1.max_depth}
    dbg = torch.zeros(1024, dtype=torch.int64, device="cuda")
    e.L.yy_engine_set_debug_stamps(e.handle, ctypes.c_void_p(dbg.data_ptr()))
    e.selfplay_advance(iters)
    torch.cuda.synchronize()
    e.L.yy_engine_set_debug_stamps(e.handle, None)
    raw = dbg.cpu().numpy()
    for name, off in (("cta0", 600), ("cta100", 760)):
        ph = raw[off:off + 128].reshape(16, 8)
        out[name] = {"phases": "tower, fc, heads+tree, barrier, zero | in fc: panel loads, MMA waits, scatter (cycles per iteration, warp 0 / warp 15)",
                     "w0": (ph[0] / iters).round(0).tolist(), "w15": (ph[15] / iters).round(0).tolist(),
                     "total_w0": float(round(ph[0, :5].sum() / iters))}
    out["mma_issuer_cta0"] = {"waited_for_weights": int(raw[920]) / iters, "waited_for_activations": int(raw[921]) / iters, "total": int(raw[922]) / iters,
                              "unit": "cycles per iteration"}
    out["flags"] = int(os.environ.get("YY_DBG_FLAGS", 0))
    print(json.dumps(out))
    e.close()
    del e
