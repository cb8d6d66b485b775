"""Multi-GPU plumbing: one process per GPU, games sharded across ranks, no data-path collective.

Self-play games are independent (self_play.py:288-335 fans games out to worker *processes* and only
collects their example lists), so every rank owns ``games_per_rank`` private trees and a replicated
network.  ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used for exactly
two things, both off the critical path:
  * broadcast of the packed bf16 weight image from rank 0 when the model changes (7.3 MB at 8x8);
  * gather of finished replay records to rank 0 (the reference's mp.Queue of example lists).
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist


def shard_games(total_games: int, world: int, rank: int):
    """Contiguous shard [lo, hi) of game indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(total_games, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank Philox key: distinct noise / sampling streams per shard, reproducible for a given world size."""
    return (seed * 0x9E3779B97F4A7C15 + rank * 0xD1B54A32D192ED03) & ((1 << 64) - 1)


def broadcast_image(image: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """Broadcast a packed weight image (uint8 tensor, same size on every rank) in place."""
    dist.broadcast(image, src=src, group=group)
    return image


def gather_records(tensors, dst=0, group=None):
    """All ranks pass a dict of equally-keyed tensors whose first dimension is the record count (may differ
    per rank).  Rank `dst` gets the concatenation over ranks (rank order), the others get None; dst=None: every
    rank gets it (the data-parallel learner's ranks all train on the global replay buffer)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    keys = sorted(tensors)
    n_local = torch.tensor([tensors[keys[0]].shape[0]], dtype=torch.int64, device=tensors[keys[0]].device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts) if counts else 0
    out = {}
    for k in keys:
        t = tensors[k].contiguous()
        row_shape, row_bytes = tuple(t.shape[1:]), t.element_size() * int(np.prod(t.shape[1:], dtype=np.int64))
        raw = t.view(torch.uint8).reshape(t.shape[0], row_bytes) if t.shape[0] else torch.zeros((0, row_bytes), dtype=torch.uint8, device=t.device)
        pad = torch.zeros((nmax, row_bytes), dtype=torch.uint8, device=t.device)   # bytes on the wire: any dtype, any backend
        pad[: raw.shape[0]] = raw
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        if dst is None or rank == dst:
            cat = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0).contiguous()
            out[k] = cat.view(t.dtype).reshape((cat.shape[0],) + row_shape)
    return out if (dst is None or rank == dst) else None


# ---- engine-level helpers (CUDA) -------------------------------------------------------------------------
def broadcast_weights(engine, src: int = 0) -> float:
    """NCCL broadcast of rank `src`'s packed weight image; every rank then points its engine at the received
    copy.  Returns seconds (device-timed)."""
    img = engine.weight_image
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    broadcast_image(img, src)
    ev1.record()
    torch.cuda.synchronize()
    engine.load_weight_image(img)
    return ev0.elapsed_time(ev1) * 1e-3


def replay_tensors(engine):
    """Device views of the engine's replay ring as int-typed tensors (what crosses NVLink)."""
    from . import _lib
    import ctypes
    st = engine.stats()
    n = min(st.examples, engine.replay_capacity)
    v = _lib.ReplayView()
    _lib.check(engine.L.yy_selfplay_replay(engine.handle, ctypes.byref(v)))
    base = engine.workspace.data_ptr()

    def view(ptr, nbytes, dtype, shape):
        return engine.workspace[ptr - base: ptr - base + nbytes].view(dtype).view(shape)
    return {"black": view(v.black, n * engine.W * 8, torch.int64, (n, engine.W)),
            "white": view(v.white, n * engine.W * 8, torch.int64, (n, engine.W)),
            "counts": view(v.counts, n * engine.A * 2, torch.int16, (n, engine.A)),
            "game_serial": view(v.game_serial, n * 4, torch.int32, (n,)),
            "ply": view(v.ply, n * 2, torch.int16, (n,)),
            "player": view(v.player, n, torch.int8, (n,))}


def gather_replay_counts(engine, dst: int = 0):
    """Gathers every rank's replay records to rank `dst` over NCCL.  Returns (seconds, records on dst)."""
    t0 = time.perf_counter()
    out = gather_records(replay_tensors(engine), dst)
    torch.cuda.synchronize()
    n = int(out["ply"].shape[0]) if out is not None else 0
    return time.perf_counter() - t0, n
