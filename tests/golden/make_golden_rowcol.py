#!/usr/bin/env python
"""Golden fixture of the optional row/column rule (YY_RULE_ROWCOL), the rule of the browser game that the Python rules
lack.  There is no JavaScript runtime in the build container, so the rule is TRANSCRIBED LITERALLY here --
checkRowColumnConstraint, src/gui/static/js/yin_yang_game.js:338-384, loop for loop (hasBlack / hasWhite / hasEmpty per
row, then per column; a line with no empty cell and only one colour violates) -- and composed exactly as isValidMove
composes it (:186-232: place the piece, connectivity, 2x2, then the row/column check on the WHOLE board, remove the
piece) on top of the UNMODIFIED Python reference's is_valid_move for the first three tests (yin_yang_logic.py:31-56,
which the JS mirrors).  Boards: random legal play, arbitrary fills, and boards built to have nearly complete lines.

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden_rowcol.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, arbitrary_board, random_play_board  # noqa: E402  (puts /root/reference on sys.path)
from src.yin_yang import YinYangGame  # noqa: E402


def check_row_column_constraint(board, rows, cols):
    """yin_yang_game.js:338-384, literally."""
    for row in range(rows):
        has_black = has_white = has_empty = False
        for col in range(cols):
            if board[row][col] == 1:
                has_black = True
            elif board[row][col] == -1:
                has_white = True
            else:
                has_empty = True
        if not has_empty and (not has_black or not has_white):
            return False
    for col in range(cols):
        has_black = has_white = has_empty = False
        for row in range(rows):
            if board[row][col] == 1:
                has_black = True
            elif board[row][col] == -1:
                has_white = True
            else:
                has_empty = True
        if not has_empty and (not has_black or not has_white):
            return False
    return True


def is_valid_move_js(logic, row, col, piece):
    """yin_yang_game.js:186-232 with the Python reference standing in for the checks the two share."""
    if not logic.is_valid_move(row, col, piece):          # on the board, empty, connectivity, 2x2 (yin_yang_logic.py:31-56)
        return False
    logic.board[row, col] = piece
    ok = check_row_column_constraint(logic.board, logic.n, logic.m)
    logic.board[row, col] = 0
    return ok


def make(n, m, seed, n_play, n_arb, n_lines):
    game = YinYangGame(n, m)
    rng = np.random.default_rng(seed)
    boards = []
    for _ in range(n_play):
        boards.append(random_play_board(game, rng, int(rng.integers(0, n * m)))[0])
    for _ in range(n_arb):
        boards.append(arbitrary_board(game, rng, float(rng.uniform(0.3, 0.98))))
    for _ in range(n_lines):                              # a line one stone short of being single-coloured, or already so
        b = arbitrary_board(game, rng, float(rng.uniform(0.1, 0.6)))
        colour = int(rng.choice([1, -1]))
        if rng.random() < 0.5:
            r = int(rng.integers(0, n)); b.board[r, :] = colour
            if rng.random() < 0.7:
                b.board[r, int(rng.integers(0, m))] = 0
        else:
            c = int(rng.integers(0, m)); b.board[:, c] = colour
            if rng.random() < 0.7:
                b.board[int(rng.integers(0, n)), c] = 0
        boards.append(b)
    arr = np.array([b.get_board() for b in boards], dtype=np.int8)
    masks = np.zeros((2, len(boards), n * m), dtype=np.uint8)
    for i, b in enumerate(boards):
        for k, piece in enumerate((1, -1)):
            for r in range(n):
                for c in range(m):
                    masks[k, i, r * m + c] = is_valid_move_js(b, r, c, piece)
    plain = np.array([[game.getValidMoves(b, p) for b in boards] for p in (1, -1)]).astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, f"rowcol_{n}x{m}.npz"), n=n, m=m, boards=arr, mask_black=masks[0], mask_white=masks[1])
    print(f"rowcol {n}x{m}: {len(boards)} boards, {int((plain != masks).sum())} of {masks.size} mask bits differ from the Python rules")


if __name__ == "__main__":
    make(4, 4, 1, 60, 60, 80)
    make(6, 6, 2, 60, 60, 80)
    make(8, 8, 3, 60, 40, 60)
    make(5, 7, 4, 30, 30, 40)
