"""Host-side mirror of src/yin_yang/ai/self_play.py: ``SelfPlayWorker`` / ``SelfPlayManager`` /
``generate_self_play_data`` with the reference's signatures and output format, driven by the batched engine
(``yy_selfplay_run``): all games of a call run concurrently on the GPU, one warp per game tree, leaves of all
games evaluated in one inference batch per simulation step.

Output (self_play.py:373-384): ``np.savez(boards=<object array of boards exposing .board / get_board()>,
policies=float64[N, A], values=float64[N])`` in ``<output_dir>/self_play_data_<unix time>.npz``.

Reference semantics kept: search runs as player 1 for every ply while the move is applied
with the real player (Q5), every example of a game receives the same value (Q6), temperature 1 for the first
``temperature_threshold`` plies then argmax with random tie-break, Dirichlet noise only at ply 0.  Not kept:
the board-aliasing bug (Q1) -- stored boards are true snapshots of the position before each move.
"""
from __future__ import annotations

import logging
import os
import time

import numpy as np

from . import engine as _engine
from .game import YinYangLogic
from .network import YinYangNeuralNetwork

logger = logging.getLogger("YinYangSelfPlay")
MAX_CONCURRENT_GAMES = int(os.environ.get("YY_MAX_CONCURRENT_GAMES", 4096))


def _load_network(game, model_path):
    net = YinYangNeuralNetwork(game)
    if os.path.exists(model_path):
        net.load_model(model_path)
        logger.info(f"Loaded model from {model_path}")
    else:
        logger.warning(f"No model found at {model_path}, using randomly initialized model")  # self_play.py:58-59
    return net


def play_games(game, state_dict, num_games, num_simulations=800, temperature_threshold=10, dirichlet_alpha=0.3,
               dirichlet_epsilon=0.25, cpuct=1.0, seed=None, evaluator="nn", search_as_black=True, max_moves=None,
               device=None):
    """Plays ``num_games`` complete self-play games concurrently.  Returns a list (one entry per game, in game
    order) of example lists ``[(YinYangLogic, pi float64[A], z float), ...]``."""
    n, m = game.getBoardSize()
    A = n * m
    slots = max(1, min(num_games, MAX_CONCURRENT_GAMES))
    if seed is None:
        seed = int(time.time_ns() & 0xFFFFFFFF)
    max_moves = max_moves or (2 * A + 4)                       # a game places at most A stones
    rounds = (num_games + slots - 1) // slots + 1
    eng = _engine.Engine(rows=n, cols=m, n_games=slots, n_sims=num_simulations, evaluator=evaluator, cpuct=cpuct,
                         rule_flags=getattr(game, "rule_flags", 0), search_as_black=search_as_black,
                         dirichlet_alpha=dirichlet_alpha, dirichlet_epsilon=dirichlet_epsilon,
                         temperature_threshold=temperature_threshold, seed=seed, state_dict=state_dict,
                         replay_capacity=slots * max_moves * rounds, device=device)
    try:
        done, moves = False, 0
        while not done and moves < max_moves * rounds:
            eng.selfplay_run(4)
            moves += 4
            if eng.stats().games_finished >= num_games:
                rp = eng.replay()
                fin = np.unique(rp["game_serial"][rp["finished"]])
                done = np.all(np.isin(np.arange(num_games), fin))
        rp = eng.replay()
    finally:
        eng.close()
    games = []
    order = np.lexsort((rp["ply"], rp["game_serial"]))
    serial_sorted = rp["game_serial"][order]
    lo = np.searchsorted(serial_sorted, np.arange(num_games), side="left")     # records of game g: order[lo[g]:hi[g]]
    hi = np.searchsorted(serial_sorted, np.arange(num_games), side="right")
    for g in range(num_games):
        idx = order[lo[g]:hi[g]]
        ex = []
        for i in idx:
            if not rp["finished"][i]:
                continue
            b = YinYangLogic(n, m, getattr(game, "rule_flags", 0))
            b.board = rp["boards"][i].copy()
            z = float(rp["z"][i])
            ex.append((b, rp["pi"][i].copy(), z if z == 0.0001 else int(z)))
        games.append(ex)
    return games


class SelfPlayWorker:
    """self_play.py:22-216.  ``play_game`` plays one game; ``generate_games`` plays ``num_games`` concurrently."""

    def __init__(self, game, model_path, num_simulations=800, num_games=1, temperature_threshold=10,
                 dirichlet_alpha=0.3, dirichlet_epsilon=0.25, cpuct=1.0, num_parallel=1):
        self.game, self.model_path = game, model_path
        self.num_simulations, self.num_games = num_simulations, num_games
        self.temperature_threshold = temperature_threshold
        self.dirichlet_alpha, self.dirichlet_epsilon, self.cpuct = dirichlet_alpha, dirichlet_epsilon, cpuct
        self.num_parallel = num_parallel
        self.neural_net = _load_network(game, model_path)

    def _play(self, count):
        return play_games(self.game, self.neural_net.state_dict(), count, self.num_simulations, self.temperature_threshold,
                          self.dirichlet_alpha, self.dirichlet_epsilon, self.cpuct)

    def play_game(self):
        return self._play(1)[0]

    def generate_games(self):
        all_examples = []
        for i, ex in enumerate(self._play(self.num_games)):
            all_examples.extend(ex)
            logger.info(f"Completed game {i + 1}/{self.num_games} with {len(ex)} examples")
        return all_examples


class SelfPlayManager:
    """self_play.py:218-335.  ``num_workers`` x ``games_per_worker`` games; the reference forks one process per
    worker -- here they all share the GPU batch (the worker count only determines how many games are played)."""

    def __init__(self, game, model_path, num_workers=1, num_simulations=800, games_per_worker=1,
                 temperature_threshold=10, dirichlet_alpha=0.3, dirichlet_epsilon=0.25, cpuct=1.0, mcts_parallel=1):
        self.game, self.model_path = game, model_path
        self.num_workers, self.num_simulations, self.games_per_worker = num_workers, num_simulations, games_per_worker
        self.temperature_threshold = temperature_threshold
        self.dirichlet_alpha, self.dirichlet_epsilon, self.cpuct = dirichlet_alpha, dirichlet_epsilon, cpuct
        self.mcts_parallel = mcts_parallel

    def generate_games_parallel(self):
        try:
            worker = SelfPlayWorker(self.game, self.model_path, self.num_simulations, self.num_workers * self.games_per_worker,
                                    self.temperature_threshold, self.dirichlet_alpha, self.dirichlet_epsilon, self.cpuct)
            examples = worker.generate_games()
        except Exception as e:  # self_play.py:283-286: a failed worker contributes no examples
            logger.error(f"Self-play failed: {e}")
            examples = []
        logger.info(f"Generated {len(examples)} examples")
        return examples


def save_examples(filename, examples, action_size, wire="native"):
    """np.savez(boards=<object array>, policies=float64[N,A], values=float64[N]) (self_play.py:373-384).

    wire="native"    : boards are this package's ``YinYangLogic`` (unpickling needs ``yy_b200`` importable);
    wire="reference" : boards are pickled as ``src.yin_yang.yin_yang_logic.YinYangLogic`` (attributes n, m, board:
                       yin_yang_logic.py:14-22), so the reference's own trainer (``TrainingDataQueue.push_file``,
                       training_pipeline.py:55-73) loads the file without this package installed."""
    boards = np.empty(len(examples), dtype=object)
    undo = None
    if wire == "reference":
        cls, undo = _reference_board_class()
        for i, ex in enumerate(examples):
            o = cls.__new__(cls)
            o.n, o.m = int(ex[0].n), int(ex[0].m)
            o.board = np.array(ex[0].get_board() if hasattr(ex[0], "get_board") else ex[0].board, dtype=np.int8)
            boards[i] = o
    elif wire == "native":
        for i, ex in enumerate(examples):
            boards[i] = ex[0]
    else:
        raise ValueError("wire must be 'native' or 'reference'")
    policies = np.array([ex[1] for ex in examples], dtype=np.float64).reshape(len(examples), action_size)
    values = np.array([ex[2] for ex in examples], dtype=np.float64)
    try:
        np.savez(filename, boards=boards, policies=policies, values=values)
    finally:
        if undo:
            import sys
            for name in undo:
                sys.modules.pop(name, None)
    return filename


def _reference_board_class():
    """(class, fake module names to remove afterwards).  pickle stores a class by module path + name and checks that the
    path resolves to the very object being pickled: use the reference's class when it is importable, otherwise register
    a stand-in under the reference's module path for the duration of the save."""
    import sys
    import types
    try:
        from src.yin_yang.yin_yang_logic import YinYangLogic as Ref   # running inside the reference tree
        return Ref, None
    except Exception:
        created = []
        for name in ("src", "src.yin_yang", "src.yin_yang.yin_yang_logic"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
                created.append(name)
        cls = type("YinYangLogic", (), {"__module__": "src.yin_yang.yin_yang_logic"})
        sys.modules["src.yin_yang.yin_yang_logic"].YinYangLogic = cls
        return cls, created


def generate_self_play_data(game, model_path, output_dir, num_games=100, num_workers=1, num_simulations=800, wire="native", file_tag=""):
    """self_play.py:337-387.  Returns the path of the written .npz (``wire``: see save_examples; ``file_tag``: suffix that keeps
    the files of several ranks apart)."""
    os.makedirs(output_dir, exist_ok=True)
    games_per_worker = max(1, num_games // num_workers)          # self_play.py:355
    manager = SelfPlayManager(game=game, model_path=model_path, num_workers=num_workers, games_per_worker=games_per_worker,
                              num_simulations=num_simulations)
    examples = manager.generate_games_parallel()
    filename = os.path.join(output_dir, f"self_play_data_{int(time.time())}{file_tag}.npz")
    save_examples(filename, examples, game.getActionSize(), wire)
    logger.info(f"Saved {len(examples)} examples to {filename}")
    return filename
