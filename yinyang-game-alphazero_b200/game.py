"""Host-side mirror of the reference's game surface, backed by the CUDA rules kernels.

``YinYangLogic`` / ``YinYangGame`` keep the names, argument meaning and return conventions of
src/yin_yang/yin_yang_logic.py and src/yin_yang/yin_yang_game.py, so code written against the reference
(``getValidMoves``, ``getNextState``, ``getGameEnded``, ...) runs unchanged.  Every rules decision is made by
``libyinyang_b200.so`` on the GPU (engine.*_host: H2D -> kernel -> D2H); nothing is decided on the CPU.

Differences, on purpose:
  * value semantics: ``getNextState`` returns a NEW board; the reference mutates its argument and returns the
    same object (yin_yang_game.py:52-58), which corrupts MCTS trees (Q1);
  * ``stringRepresentation`` works on numpy >= 2 (the reference's ``tostring()`` raises, Q9);
  * batched variants (``*_batch``) take int8[B, n, m] arrays -- one kernel launch for B boards.
``rule_flags=RULE_ROWCOL`` additionally applies the browser game's row/column rule
(src/gui/static/js/yin_yang_game.js:338-384); the default (0) is the Python rule set used by self-play.
"""
from __future__ import annotations

import numpy as np

from . import engine as _engine

RULE_ROWCOL = _engine.RULE_ROWCOL


class YinYangLogic:
    """Board container (yin_yang_logic.py:4-134): int8[n, m], 0 empty / +1 black / -1 white."""

    def __init__(self, n=8, m=8, rule_flags=0):
        self.n, self.m, self.rule_flags = n, m, rule_flags
        self.board = np.zeros((n, m), dtype=np.int8)

    def get_board(self):
        return self.board.copy()

    def copy(self):
        b = YinYangLogic(self.n, self.m, self.rule_flags)
        b.board = self.board.copy()
        return b

    def _mask(self, piece):
        return _engine.legal_mask_host(self.board[None], np.array([1 if piece == 1 else -1], np.int8), self.n, self.m,
                                       self.rule_flags)[0]

    def is_valid_move(self, x, y, piece):  # yin_yang_logic.py:31-56
        if not (0 <= x < self.n and 0 <= y < self.m):
            return False
        return bool(self._mask(piece)[x * self.m + y])

    def place_piece(self, x, y, piece):  # yin_yang_logic.py:24-29 (in place, like the reference)
        if self.is_valid_move(x, y, piece):
            self.board[x, y] = piece
            return True
        return False

    def get_valid_moves(self, piece):  # yin_yang_logic.py:111-120
        return [(int(a) // self.m, int(a) % self.m) for a in np.flatnonzero(self._mask(piece))]

    def has_valid_move(self, piece):  # yin_yang_logic.py:122-128
        return bool(self._mask(piece).any())

    def count_pieces(self):  # yin_yang_logic.py:130-134
        return np.sum(self.board == 1), np.sum(self.board == -1)


class YinYangGame:
    """alpha-zero-general style Game API (yin_yang_game.py:10-207)."""

    def __init__(self, n=8, m=8, rule_flags=0):
        self.n, self.m, self.rule_flags = n, m, rule_flags
        self.action_size = n * m

    # -- reference surface
    def getInitBoard(self):
        return YinYangLogic(self.n, self.m, self.rule_flags)

    def getBoardSize(self):
        return (self.n, self.m)

    def getActionSize(self):
        return self.action_size

    def getNextState(self, board, player, action):  # yin_yang_game.py:39-58 (illegal action: silent no-op)
        nb, npl = _engine.next_state_host(board.board[None], np.array([player], np.int8), np.array([action], np.int32),
                                          self.n, self.m, self.rule_flags)
        out = YinYangLogic(self.n, self.m, self.rule_flags)
        out.board = nb[0]
        return out, -player

    def getValidMoves(self, board, player):  # yin_yang_game.py:60-78 -> float64[A] of 0/1
        return self.getValidMovesBatch(board.board[None], [player])[0]

    def getGameEnded(self, board, player):  # yin_yang_game.py:80-110 -> 0 | 1 | -1 | 0.0001
        r = float(self.getGameEndedBatch(board.board[None], [player])[0])
        return r if r == 0.0001 else int(r)

    def getCanonicalForm(self, board, player):  # yin_yang_game.py:112-125: identity
        return board

    def getSymmetries(self, board, pi):  # yin_yang_game.py:127-166 (square boards, like the reference)
        pi_board = np.reshape(pi, (self.n, self.m))
        b = board.get_board()
        syms = []
        for i in range(1, 5):
            for flip in (True, False):
                nb, npi = np.rot90(b, i), np.rot90(pi_board, i)
                if flip:
                    nb, npi = np.fliplr(nb), np.fliplr(npi)
                sb = YinYangLogic(nb.shape[0], nb.shape[1], self.rule_flags)
                sb.board = np.ascontiguousarray(nb)
                syms.append((sb, npi.flatten()))
        return syms

    def stringRepresentation(self, board):  # yin_yang_game.py:168-178
        return board.get_board().tobytes()

    def _action_to_coords(self, action):
        return action // self.m, action % self.m

    def _coords_to_action(self, x, y):
        return x * self.m + y

    def display(self, board):  # yin_yang_game.py:188-207
        b = board.get_board()
        print(" " + "".join(chr(97 + i) for i in range(self.m)))
        for i in range(self.n):
            print(str(i + 1) + "".join("B" if v == 1 else "W" if v == -1 else "." for v in b[i]))

    # -- batched variants (one kernel launch for B boards)
    def getValidMovesBatch(self, boards, players):
        m = _engine.legal_mask_host(boards, np.asarray(players, np.int8), self.n, self.m, self.rule_flags)
        return m.astype(np.float64)

    def getNextStateBatch(self, boards, players, actions):
        return _engine.next_state_host(boards, np.asarray(players, np.int8), np.asarray(actions, np.int32), self.n, self.m,
                                       self.rule_flags)

    def getGameEndedBatch(self, boards, players):
        return _engine.ended_host(boards, np.asarray(players, np.int8), self.n, self.m, self.rule_flags)

    def envStepBatch(self, boards, players, actions):
        """mask of the side to move, successor boards/players, terminal value of the successor."""
        return _engine.env_step_host(boards, np.asarray(players, np.int8), np.asarray(actions, np.int32), self.n, self.m,
                                     self.rule_flags)
