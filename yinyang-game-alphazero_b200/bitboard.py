"""Host-side packing between the reference's int8 boards and the engine's bitboards.

Reference board: ``np.int8[n, m]`` with 0 empty / +1 black / -1 white
(src/yin_yang/yin_yang_logic.py:14-22).  Engine board: W = ceil(n*m/64) little-endian
64-bit words per colour, bit ``a = x*m + y`` (the action index, yin_yang_game.py:180-186).
"""
from __future__ import annotations

import numpy as np


def words_for(n: int, m: int) -> int:
    return (n * m + 63) // 64


def pack_boards(boards, n: int, m: int):
    """int8[B, n, m] (or [B, n*m]) -> (black uint64[B, W], white uint64[B, W])."""
    b = np.ascontiguousarray(boards, dtype=np.int8).reshape(-1, n * m)
    W = words_for(n, m)
    pad = W * 64 - n * m
    out = []
    for colour in (1, -1):
        bits = (b == colour).astype(np.uint8)
        if pad:
            bits = np.concatenate([bits, np.zeros((bits.shape[0], pad), np.uint8)], axis=1)
        packed = np.packbits(bits.reshape(-1, W, 64), axis=-1, bitorder="little")
        out.append(np.ascontiguousarray(packed).view(np.uint64).reshape(-1, W))
    return out[0], out[1]


def unpack_bits(words, n: int, m: int) -> np.ndarray:
    """uint64[B, W] -> uint8[B, n*m] of 0/1."""
    w = np.ascontiguousarray(words, dtype=np.uint64)
    W = words_for(n, m)
    w = w.reshape(-1, W)
    bits = np.unpackbits(w.view(np.uint8).reshape(-1, W * 8), axis=-1, bitorder="little")
    return np.ascontiguousarray(bits[:, : n * m])


def unpack_boards(black, white, n: int, m: int) -> np.ndarray:
    """(uint64[B, W], uint64[B, W]) -> int8[B, n, m]."""
    bb = unpack_bits(black, n, m).astype(np.int8)
    ww = unpack_bits(white, n, m).astype(np.int8)
    return (bb - ww).reshape(-1, n, m)
