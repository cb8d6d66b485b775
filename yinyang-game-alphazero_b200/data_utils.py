"""Host-side mirror of src/yin_yang/ai/data_utils.py: replay examples -> training tensors, on the GPU.

``create_dataset_from_games(game_data, game, augment=True)`` keeps the reference's signature and return value (three
lists of torch tensors, one entry per sample: planes [5,n,m], policy [A], value [1]; data_utils.py:182-215) but runs
preprocess_sample + augment_sample for the whole list in ONE kernel (csrc/yy_dataset.cu, yy_augment_samples).
``augment_replay(engine)`` does the same straight from an engine's replay ring without leaving the device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import bitboard
from . import engine as _engine


class DataProcessor:
    """data_utils.py:7-180.  preprocess_sample / augment_sample for single samples (thin wrappers over the batch path)."""

    def __init__(self, game):
        self.game = game
        self.board_size = game.getBoardSize()

    def preprocess_sample(self, board, policy, player):
        n, m = self.board_size
        planes, pol, _ = _engine.augment_samples_host(_board_array(board)[None], n, m, policy=np.asarray(policy, np.float32)[None], forms=1)
        return torch.from_numpy(planes[0]), torch.from_numpy(pol[0])

    def augment_sample(self, board, policy):
        """board: the plane tensor of preprocess_sample (as in the reference), or a board."""
        n, m = self.board_size
        planes, pol, _ = _engine.augment_samples_host(_board_array(board)[None], n, m, policy=np.asarray(policy, np.float32)[None])
        return [(torch.from_numpy(planes[f]), torch.from_numpy(pol[f])) for f in range(8)]


def _board_array(board):
    """int8 board from a YinYangLogic-like object, an int8 array, or the reference's plane tensor [5,n,m]
    (board_to_input output: channel 1 = black, channel 2 = white; the planes are recomputed on the device)."""
    if hasattr(board, "get_board"):
        return np.asarray(board.get_board(), dtype=np.int8)
    if isinstance(board, torch.Tensor):
        board = board.detach().cpu().numpy()
    board = np.asarray(board)
    if board.ndim == 3 and board.shape[0] == 5:
        return (board[1] - board[2]).astype(np.int8)
    return board.astype(np.int8)


def dataset_tensors(game_data, game, augment=True):
    """(planes float32[S,5,n,m], policies float32[S,A], values float32[S]) device tensors, S = 8 * len(game_data) with
    augmentation (sample 8r+f = form f of example r), else len(game_data)."""
    n, m = game.getBoardSize()
    boards = np.stack([_board_array(b) for b, _, _ in game_data]) if len(game_data) else np.zeros((0, n, m), np.int8)
    pol = np.stack([np.asarray(p, dtype=np.float64) for _, p, _ in game_data]).astype(np.float32) if len(game_data) else np.zeros((0, n * m), np.float32)
    val = np.asarray([v for _, _, v in game_data], dtype=np.float64).astype(np.float32)
    b, w = bitboard.pack_boards(boards, n, m)
    bd, wd = _engine._to_dev(b, torch.int64), _engine._to_dev(w, torch.int64)
    return _engine.augment_samples(bd, wd, n, m, policy=_engine._to_dev(pol, torch.float32),
                                   values=_engine._to_dev(val, torch.float32), forms=8 if augment else 1)


def create_dataset_from_games(game_data, game, augment=True):
    """data_utils.py:182-215: lists of per-sample tensors (board [5,n,m], policy [A], value [1])."""
    planes, policies, values = dataset_tensors(game_data, game, augment)
    planes, policies, values = planes.cpu(), policies.cpu(), values.cpu()
    return list(planes.unbind(0)), list(policies.unbind(0)), [v.reshape(1) for v in values.unbind(0)]


def augment_replay(eng, finished_only=True):
    """Training tensors straight from an Engine's replay ring (device in, device out).  Values: the game result from the
    record's perspective (self_play.py:171-180: the same sign convention as Engine.replay()['z'])."""
    rp = eng.replay()
    keep = rp["finished"] if finished_only else np.ones(len(rp["z"]), bool)
    boards, counts, z = rp["boards"][keep], rp["counts"][keep], np.nan_to_num(rp["z"][keep])
    b, w = bitboard.pack_boards(boards, eng.rows, eng.cols)
    return _engine.augment_samples(_engine._to_dev(b, torch.int64), _engine._to_dev(w, torch.int64), eng.rows, eng.cols,
                                   counts=_engine._to_dev(counts.view(np.int16), torch.int16),
                                   values=_engine._to_dev(z.astype(np.float32), torch.float32))
