"""Python/numpy/torch port of the reference's CPU self-play path -- TEST INFRASTRUCTURE ONLY.

This is what ``bench.py``'s ``cpu_baseline`` and ``--impl reference`` legs time (kind "port"): the upstream
reference is pure Python and cannot travel to the GPU box, so its algorithm is restated here with the SAME
structure and cost profile -- per-cell trial placement + BFS legality on an int8 numpy board, a tree of
Python node objects searched sequentially, one batch-1 fp32 torch forward per leaf on the host CPU.
It is pinned against tests/golden/ (produced by the unmodified reference) in tests/test_oracle_golden.py.

Citations are to the upstream repository: rules src/yin_yang/yin_yang_logic.py, adapter
src/yin_yang/yin_yang_game.py, search src/yin_yang/ai/mcts.py, network src/yin_yang/ai/neural_network.py,
episode driver src/yin_yang/ai/self_play.py.
"""
from __future__ import annotations

import math
from collections import deque

import numpy as np

NEIGH = ((0, 1), (1, 0), (0, -1), (-1, 0))


# ------------------------------------------------------------------------------------------------ rules
class Board:
    """int8[n, m] grid, 0 empty / +1 black / -1 white (yin_yang_logic.py:14-22)."""

    def __init__(self, n, m, grid=None):
        self.n, self.m = n, m
        self.board = np.zeros((n, m), dtype=np.int8) if grid is None else np.array(grid, dtype=np.int8)

    def get_board(self):
        return self.board.copy()

    def clone(self):
        return Board(self.n, self.m, self.board)

    def _one_component(self, colour):  # yin_yang_logic.py:58-94
        cells = np.argwhere(self.board == colour)
        if len(cells) == 0:
            return True
        start = (int(cells[0][0]), int(cells[0][1]))
        seen = {start}
        todo = deque([start])
        while todo:
            r, c = todo.popleft()
            for dr, dc in NEIGH:
                q = (r + dr, c + dc)
                if 0 <= q[0] < self.n and 0 <= q[1] < self.m and q not in seen and self.board[q] == colour:
                    seen.add(q)
                    todo.append(q)
        return len(seen) == len(cells)

    def _no_square(self):  # yin_yang_logic.py:96-109
        g = self.board
        for r in range(self.n - 1):
            for c in range(self.m - 1):
                v = g[r, c]
                if v != 0 and g[r, c + 1] == v and g[r + 1, c] == v and g[r + 1, c + 1] == v:
                    return False
        return True

    def legal(self, r, c, colour):  # yin_yang_logic.py:31-56
        if not (0 <= r < self.n and 0 <= c < self.m) or self.board[r, c] != 0:
            return False
        self.board[r, c] = colour
        ok = self._one_component(colour) and self._no_square()
        self.board[r, c] = 0
        return ok

    def legal_cells(self, colour):  # yin_yang_logic.py:111-120
        return [(r, c) for r in range(self.n) for c in range(self.m) if self.legal(r, c, colour)]

    def can_move(self, colour):  # yin_yang_logic.py:122-128
        return any(self.legal(r, c, colour) for r in range(self.n) for c in range(self.m))


class Game:
    """Value-semantics restatement of YinYangGame (yin_yang_game.py:10-110)."""

    def __init__(self, n=8, m=8):
        self.n, self.m, self.action_size = n, m, n * m

    def getInitBoard(self):
        return Board(self.n, self.m)

    def getActionSize(self):
        return self.action_size

    def getBoardSize(self):
        return (self.n, self.m)

    def getValidMoves(self, board, player):  # :60-78
        out = np.zeros(self.action_size)
        for r, c in board.legal_cells(1 if player == 1 else -1):
            out[r * self.m + c] = 1
        return out

    def getNextState(self, board, player, action):  # :39-58, on a copy
        nb = board.clone()
        r, c = divmod(int(action), self.m)
        colour = 1 if player == 1 else -1
        if nb.legal(r, c, colour):
            nb.board[r, c] = colour
        return nb, -player

    def getGameEnded(self, board, player):  # :80-110
        me = 1 if player == 1 else -1
        if board.can_move(me) or board.can_move(-me):
            return 0
        black, white = int((board.board == 1).sum()), int((board.board == -1).sum())
        if black > white:
            return 1 if player == 1 else -1
        if white > black:
            return 1 if player == -1 else -1
        return 0.0001


# ------------------------------------------------------------------------------------------------ search
class TreeNode:  # mcts.py:28-48
    __slots__ = ("prior", "visits", "value_sum", "children", "board", "player", "terminal", "terminal_value", "action")

    def __init__(self, prior=0.0, action=None):
        self.prior, self.action = prior, action
        self.visits, self.value_sum = 0, 0.0
        self.children = {}
        self.board = self.player = None
        self.terminal, self.terminal_value = False, None


def expand(game, node, board, player, policy):  # mcts.py:50-91
    node.board, node.player = board, player
    res = game.getGameEnded(board, player)
    if res != 0:
        node.terminal, node.terminal_value = True, res
        return
    for a in np.flatnonzero(game.getValidMoves(board, player) == 1):
        node.children[int(a)] = TreeNode(policy[a], int(a))


def pick_child(node, cpuct):  # mcts.py:97-145 (strict '>' => first maximum in ascending action order)
    total = sum(ch.visits for ch in node.children.values())
    best, best_ucb = None, -float("inf")
    for ch in node.children.values():
        q = ch.value_sum / ch.visits if ch.visits > 0 else 0.0
        ucb = q + cpuct * ch.prior * math.sqrt(total) / (1 + ch.visits)
        if ucb > best_ucb:
            best, best_ucb = ch, ucb
    return best


def simulate(game, net, root, cpuct):  # mcts.py:345-414
    path, cur = [root], root
    while (cur.children or cur.terminal) and not cur.terminal:
        cur = pick_child(cur, cpuct)
        path.append(cur)
    if cur.terminal:
        value = cur.terminal_value
    elif len(path) < 2:
        policy, value = net.predict(root.board)
        expand(game, root, root.board, root.player, policy)
    else:
        parent = path[-2]
        nb, npl = game.getNextState(parent.board, parent.player, cur.action)
        policy, value = net.predict(nb)
        expand(game, cur, nb, npl, policy)
    leaf_player = path[-1].player
    for nd in reversed(path):  # mcts.py:406-412
        nd.visits += 1
        nd.value_sum += -value if (nd is not path[-1] and nd.player != leaf_player) else value


def search(game, net, board, player, num_sims, cpuct=1.0, noise=None, eps=0.25):  # mcts.py:275-343
    root = TreeNode()
    root.board, root.player = board, player
    policy, _ = net.predict(board)
    if noise is not None:
        idx = np.flatnonzero(game.getValidMoves(board, player) == 1)
        for i, a in enumerate(idx):
            policy[a] = (1 - eps) * policy[a] + eps * noise[i]
    expand(game, root, board, player, policy)
    for _ in range(num_sims):
        simulate(game, net, root, cpuct)
    counts = np.zeros(game.getActionSize())
    for a, ch in root.children.items():
        counts[a] = ch.visits
    return counts, root


# ------------------------------------------------------------------------------------------------ network
def build_net(n, m, channels=128, blocks=10):
    """fp32 torch restatement of YinYangNeuralNetwork (neural_network.py:16-123) with identical state_dict keys."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    class Block(nn.Module):  # neural_network.py:16-33
        def __init__(self, c):
            super().__init__()
            self.conv1, self.bn1 = nn.Conv2d(c, c, 3, padding=1), nn.BatchNorm2d(c)
            self.conv2, self.bn2 = nn.Conv2d(c, c, 3, padding=1), nn.BatchNorm2d(c)

        def forward(self, x):
            y = F.relu(self.bn1(self.conv1(x)))
            return F.relu(self.bn2(self.conv2(y)) + x)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            A = n * m
            self.n, self.m, self.A = n, m, A
            self.conv1, self.bn1 = nn.Conv2d(5, channels, 3, padding=1), nn.BatchNorm2d(channels)
            self.res_blocks = nn.ModuleList([Block(channels) for _ in range(blocks)])
            self.policy_conv, self.policy_bn = nn.Conv2d(channels, 32, 1), nn.BatchNorm2d(32)
            self.policy_fc = nn.Linear(32 * A, A)
            self.value_conv, self.value_bn = nn.Conv2d(channels, 32, 1), nn.BatchNorm2d(32)
            self.value_fc1, self.value_fc2 = nn.Linear(32 * A, 256), nn.Linear(256, 1)
            for mod in self.modules():  # neural_network.py:85-92
                if isinstance(mod, (nn.Conv2d, nn.Linear)):
                    nn.init.xavier_normal_(mod.weight)
                    nn.init.zeros_(mod.bias)

        def forward(self, x):  # neural_network.py:94-123
            x = F.relu(self.bn1(self.conv1(x)))
            for blk in self.res_blocks:
                x = blk(x)
            p = F.relu(self.policy_bn(self.policy_conv(x))).reshape(-1, 32 * self.A)
            v = F.relu(self.value_bn(self.value_conv(x))).reshape(-1, 32 * self.A)
            return self.policy_fc(p), torch.tanh(self.value_fc2(F.relu(self.value_fc1(v))))

        def planes(self, grids):  # board_to_input, neural_network.py:156-196, batched: int8[B,n,m] -> f32[B,5,n,m]
            g = np.asarray(grids, dtype=np.int8).reshape(-1, self.n, self.m)
            x = np.zeros((g.shape[0], 5, self.n, self.m), dtype=np.float32)
            x[:, 0], x[:, 1], x[:, 2] = g == 0, g == 1, g == -1
            filled = g != 0
            x[:, 3] = (filled.sum(axis=2, keepdims=True) / self.m).astype(np.float32)
            x[:, 4] = (filled.sum(axis=1, keepdims=True) / self.n).astype(np.float32)
            return torch.from_numpy(x)

        def predict(self, board):  # neural_network.py:125-154: eval, batch 1, softmax over all A
            self.eval()
            with torch.no_grad():
                logits, v = self.forward(self.planes(board.get_board()))
                return F.softmax(logits, dim=1).numpy()[0], v.numpy()[0][0]

    return Net()


class HashStubNet:
    """predict() backed by the C oracle's deterministic evaluator (same spec as the CUDA stub mode)."""

    def __init__(self, n, m):
        self.n, self.m = n, m

    def predict(self, board):
        import oracle
        return oracle.stub_predict(board.get_board(), self.n, self.m)


# ------------------------------------------------------------------------------------------------ timed legs
def random_play_position(game, plies, seed=0):
    """The board after `plies` uniformly random legal plies from the empty board (a side without a move passes)."""
    rng = np.random.default_rng(seed)
    board, player = game.getInitBoard(), 1
    for _ in range(plies):
        idx = np.flatnonzero(game.getValidMoves(board, player))
        if idx.size == 0:
            player = -player
            continue
        board, player = game.getNextState(board, player, int(rng.choice(idx)))
    return board


def time_searches(n, m, num_sims, n_searches, channels=128, blocks=10, threads=1, seed=0, plies=0):
    """Wall-clock seconds for n_searches full searches (one search = one self-play move) of the position after `plies`
    random plies, searched as player 1 like the reference's self-play (self_play.py:135); plies = 0: the empty board."""
    import time
    import torch
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    net = build_net(n, m, channels, blocks)
    game = Game(n, m)
    root = random_play_position(game, plies, seed)
    t0 = time.perf_counter()
    for _ in range(n_searches):
        search(game, net, root, 1, num_sims)
    return time.perf_counter() - t0


def _worker(args):
    return time_searches(*args)


def time_searches_parallel(n, m, num_sims, n_searches_per_worker, workers, channels=128, blocks=10):
    """The reference's only scaling axis: one process per worker (self_play.py:288-335), 1 torch thread each.
    Returns (total searches, wall seconds)."""
    import multiprocessing as mp
    import time
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        pool.map(_worker, [(n, m, num_sims, n_searches_per_worker, channels, blocks, 1, i) for i in range(workers)])
    return workers * n_searches_per_worker, time.perf_counter() - t0


def _env_worker(args):
    import time
    n, m, steps, seed = args
    t0 = time.perf_counter()
    env_steps(n, m, steps, seed)
    return time.perf_counter() - t0


def env_steps(n, m, n_steps, seed=0):
    """Random-play env loop (mask + step + ended), the BASELINE.md env-steps/s unit.  Returns steps done."""
    rng = np.random.default_rng(seed)
    game = Game(n, m)
    board, player, done = game.getInitBoard(), 1, 0
    while done < n_steps:
        mask = game.getValidMoves(board, player)
        idx = np.flatnonzero(mask)
        a = int(rng.choice(idx)) if idx.size else 0
        board, player = game.getNextState(board, player, a)
        res = game.getGameEnded(board, player)
        done += 1
        if res != 0:
            board, player = game.getInitBoard(), 1
    return done


def training_step(net, optimizer, planes, policies, values):
    """One optimisation step exactly as the inner loop of AlphaZeroTrainer.train runs it (src/yin_yang/ai/trainer.py:120-137):
    train() mode forward, CrossEntropyLoss with probability targets + MSELoss, backward, optimizer.step().  Returns
    (policy_loss, value_loss, {name: gradient}).  TEST INFRASTRUCTURE (the checker of the CUDA learner step)."""
    import torch
    net.train()
    optimizer.zero_grad()
    policy_logits, value_preds = net(planes)
    policy_loss = torch.nn.CrossEntropyLoss()(policy_logits, policies)
    value_loss = torch.nn.MSELoss()(value_preds.view(-1), values)
    (policy_loss + value_loss).backward()
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    optimizer.step()
    return policy_loss.item(), value_loss.item(), grads

