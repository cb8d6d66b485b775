"""yinyang-game-alphazero_b200 -- B200-native batched AlphaZero self-play engine for the Yin-Yang game.

A drop-in for the self-play hot path of Arash-san/YinYang-Game-AlphaZero (rules -> MCTS -> policy/value
inference), written from scratch as hand-written sm_100a CUDA kernels behind a C ABI
(``include/yinyang_b200.h``, ``lib/libyinyang_b200.so``).  The directory name contains a hyphen, so import
it through the root-level shim::

    import yy_b200                      # registers this package as ``yinyang_game_alphazero_b200``
    from yinyang_game_alphazero_b200 import Engine, YinYangGame, MCTS

Layout
  csrc/        CUDA kernels + the C ABI (rules bitboards, HBM tree, tcgen05 inference)
  _lib.py      nvcc build + ctypes binding
  engine.py    device/host buffer API over the C ABI
  weights.py   reference checkpoint -> packed bf16 weight image (BN folded)
  game.py, network.py, mcts.py, self_play.py, players.py, data_utils.py, trainer.py   host-side mirror of the reference's Python interface
  learner.py   one optimisation step of the reference trainer as a CUDA graph over the yy_lrn_* kernels (csrc/yy_learn.cu)
"""
from . import _lib, bitboard  # noqa: F401
from ._lib import YinYangError, build  # noqa: F401

__all__ = ["build", "YinYangError", "bitboard"]


_MODULES = ("engine", "game", "network", "mcts", "self_play", "players", "weights", "distributed", "data_utils", "arena",
            "learner", "trainer", "training_pipeline", "alphazero", "ai_move")


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `build()` works in a bare interpreter
    import importlib
    if name in _MODULES:                       # `from . import engine` inside the package: import just that module
        return importlib.import_module(f"{__name__}.{name}")
    if name.startswith("__"):
        raise AttributeError(name)
    for mod in _MODULES:                       # `from yinyang_game_alphazero_b200 import Engine, YinYangGame, MCTS`
        m = importlib.import_module(f"{__name__}.{mod}")
        if hasattr(m, name):
            return getattr(m, name)
    raise AttributeError(name)
