"""GPU tests of the reference-facing Python interface: they read like the reference's own usage
(YinYangGame / MCTS.search / generate_self_play_data / train_alphazero.py modes)."""
import os

import numpy as np
import pytest

from conftest import golden_files, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(yy):
    from yinyang_game_alphazero_b200 import game, mcts, network, players, self_play
    return dict(game=game, mcts=mcts, network=network, players=players, self_play=self_play)


def test_game_surface_matches_reference_golden(pkg):
    g = load_golden("rules_6x6.npz")
    game = pkg["game"].YinYangGame(6, 6)
    assert game.getBoardSize() == (6, 6) and game.getActionSize() == 36
    for i in range(0, len(g["boards"]), 40):
        b = game.getInitBoard()
        b.board = g["boards"][i].copy()
        vm = game.getValidMoves(b, 1)
        assert vm.dtype == np.float64 and np.array_equal(vm.astype(np.uint8), g["mask_black"][i])
        nb, npl = game.getNextState(b, int(g["players"][i]), int(g["actions"][i]))
        assert np.array_equal(nb.board, g["next_boards"][i]) and npl == g["next_players"][i]
        assert np.array_equal(b.board, g["boards"][i])                       # value semantics: input untouched
        r = game.getGameEnded(b, -1)
        assert r == g["ended_white"][i] and (isinstance(r, int) or r == 0.0001)
        assert b.has_valid_move(1) == bool(g["mask_black"][i].any())
        assert sorted(b.get_valid_moves(-1)) == [(a // 6, a % 6) for a in np.flatnonzero(g["mask_white"][i])]
    e = game.getInitBoard()
    assert e.place_piece(2, 3, 1) and e.board[2, 3] == 1 and not e.place_piece(2, 3, -1)
    assert game._coords_to_action(*game._action_to_coords(17)) == 17
    assert len(game.getSymmetries(e, np.arange(36) / 630.0)) == 8
    assert isinstance(game.stringRepresentation(e), bytes)


@pytest.mark.parametrize("name", ["mcts_6x6_s400.npz", "mcts_4x4_s300.npz", "mcts_8x8_s200_mid_white.npz"])
def test_mcts_facade_matches_reference_golden(pkg, name):
    """MCTS(game, evaluator, num_simulations=...).search(board, player) -> (probs, root), like mcts_tests.py."""
    g = load_golden(name)
    n, m = int(g["n"]), int(g["m"])
    game = pkg["game"].YinYangGame(n, m)
    board = game.getInitBoard()
    board.board = g["board"].copy()
    mcts = pkg["mcts"].MCTS(game, pkg["network"].HashStubEvaluator(game), num_simulations=int(g["sims"]),
                            cpuct=float(g["cpuct"]), dirichlet_noise=False, verbose=0)
    probs, root = mcts.search(board, int(g["player"]))
    assert np.array_equal(root.get_children_visit_counts().astype(np.int32), g["counts"])
    assert probs.dtype == np.float64 and np.array_equal(probs, g["probs"])   # float64 counts / sum (mcts.py:209)
    assert abs(probs.sum() - 1.0) < 1e-12 and root.visits == int(g["root_visits"])
    for a, ch in root.children.items():
        assert ch.visits == g["counts"][a] and ch.value_sum == g["child_w"][a]
    assert np.array_equal(board.board, g["board"])
    mcts.close()


def test_mcts_external_python_evaluator(pkg, oracle_mod):
    """Any object with predict(board) plugs in (the reference's duck-typed seam, mcts_tests.py:22-32)."""
    class PyStub:
        calls = 0

        def predict(self, board):
            PyStub.calls += 1
            return oracle_mod.stub_predict(board.get_board(), 6, 6)
    game = pkg["game"].YinYangGame(6, 6)
    mcts = pkg["mcts"].MCTS(game, PyStub(), num_simulations=50, dirichlet_noise=False, verbose=0)
    probs, root = mcts.search(game.getInitBoard(), 1)
    r = oracle_mod.mcts_search(np.zeros((6, 6), np.int8), 1, 6, 6, 50)
    assert np.array_equal(root.get_children_visit_counts().astype(np.int32), r["counts"]) and PyStub.calls == 51
    a = mcts.select_action(game.getInitBoard(), 1, temperature=0)
    assert a == int(np.argmax(r["counts"]))
    mcts.close()


def test_network_predict_contract(pkg):
    import torch
    game = pkg["game"].YinYangGame(6, 6)
    torch.manual_seed(0)
    net = pkg["network"].YinYangNeuralNetwork(game, num_channels=128, num_res_blocks=2)
    p, v = net.predict(game.getInitBoard())
    assert p.shape == (36,) and p.dtype == np.float32 and abs(p.sum() - 1) < 1e-5 and -1 <= float(v) <= 1
    assert isinstance(v, np.float32)


def test_generate_self_play_data_and_cli(pkg, tmp_path):
    import torch
    import train_alphazero
    game = pkg["game"].YinYangGame(4, 4)
    model_dir, data_dir = tmp_path / "models", tmp_path / "data"
    torch.manual_seed(0)
    pkg["network"].YinYangNeuralNetwork(game).save_model(str(model_dir / "best_model.pth.tar"))
    path = pkg["self_play"].generate_self_play_data(game, str(model_dir / "best_model.pth.tar"), str(data_dir),
                                                    num_games=6, num_workers=2, num_simulations=16)
    d = np.load(path, allow_pickle=True)
    assert set(d.files) == {"boards", "policies", "values"}                     # self_play.py:379-384
    N = len(d["values"])
    assert N >= 6 and d["policies"].shape == (N, 16) and d["policies"].dtype == np.float64
    np.testing.assert_allclose(d["policies"].sum(axis=1), 1.0, atol=1e-12)
    assert np.all(np.isin(d["values"], [1.0, -1.0, 0.0001]))
    assert d["boards"][0].get_board().shape == (4, 4) and hasattr(d["boards"][0], "board")
    assert np.abs(d["boards"][0].board).sum() == 0                              # first example of game 0: empty board
    # CLI modes
    assert train_alphazero.main(["--mode", "self-play", "--rows", "4", "--cols", "4", "--simulations", "8", "--episodes", "3",
                                 "--model-dir", str(model_dir), "--data-dir", str(data_dir)]) == 0
    assert train_alphazero.main(["--mode", "evaluate", "--rows", "4", "--cols", "4", "--simulations", "8", "--eval-games", "2",
                                 "--model-dir", str(model_dir), "--data-dir", str(data_dir)]) == 0
    assert train_alphazero.main(["--mode", "self-play", "--model-dir", str(tmp_path / "none"), "--data-dir", str(data_dir)]) == 1
    assert len([f for f in os.listdir(data_dir) if f.endswith(".npz")]) >= 1


def test_data_utils_facade_matches_reference_golden(pkg):
    """create_dataset_from_games / DataProcessor with the reference's signatures (data_utils.py:7-215)."""
    import torch
    from conftest import load_golden
    from yinyang_game_alphazero_b200 import data_utils
    g = load_golden("augment_6x6.npz")
    n, m = int(g["n"]), int(g["m"])
    game = pkg["game"].YinYangGame(n, m)
    data = []
    for b, c, z in zip(g["boards"], g["counts"], g["values"]):
        tot = float(c.sum())
        pi = c.astype(np.float64) / tot if tot > 0 else np.ones(n * m) / (n * m)
        board = game.getInitBoard()
        board.board[:] = b
        data.append((board, pi, float(z)))
    bt, pt, vt = data_utils.create_dataset_from_games(data, game, augment=True)
    assert len(bt) == len(pt) == len(vt) == 8 * len(data)
    assert bt[0].shape == (5, n, m) and pt[0].shape == (n * m,) and vt[0].shape == (1,)
    assert np.array_equal(torch.stack(bt).numpy(), g["planes"])
    assert np.array_equal(torch.stack(pt).numpy(), g["policies"])
    assert np.array_equal(torch.cat(vt).numpy(), g["out_values"])
    bt1, pt1, vt1 = data_utils.create_dataset_from_games(data, game, augment=False)
    assert len(bt1) == len(data) and np.array_equal(torch.stack(bt1).numpy(), g["planes"][0::8])
    proc = data_utils.DataProcessor(game)
    x, p = proc.preprocess_sample(data[0][0], data[0][1], 1)
    forms = proc.augment_sample(x, p)                       # the reference passes the plane tensor here
    assert len(forms) == 8 and np.array_equal(forms[3][0].numpy(), g["planes"][3]) and np.array_equal(forms[3][1].numpy(), g["policies"][3])


def test_arena_matches_sequential_reference_loop(pkg):
    """arena.play_match (all games in lock-step, two batched searches per ply) against the reference's own loop
    (alphazero.py:176-219) written with the facade's MCTS.select_action / game.getNextState / getGameEnded, which are
    pinned against the reference goldens.  Deterministic evaluators; the two 'models' differ in c_puct."""
    from yinyang_game_alphazero_b200 import arena
    n = m = 5
    game = pkg["game"].YinYangGame(n, m)
    stub = pkg["network"].HashStubEvaluator(game)
    cur = pkg["mcts"].MCTS(game, stub, num_simulations=40, cpuct=1.0, dirichlet_noise=False, verbose=0)
    best = pkg["mcts"].MCTS(game, stub, num_simulations=40, cpuct=3.0, dirichlet_noise=False, verbose=0)
    G = 6
    out = arena.play_match(game, cur, best, G)
    cw = bw = dr = 0
    results = []
    for i in range(G):
        first, second, first_is_current = (cur, best, True) if i % 2 == 0 else (best, cur, False)
        board, player = game.getInitBoard(), 1
        for _ in range(4 * n * m + 8):
            action = (first if player == 1 else second).select_action(board, player, temperature=0)
            board, player = game.getNextState(board, player, action)
            r = game.getGameEnded(board, player)
            if r != 0:
                break
        results.append(r)
        if r == 1:
            cw, bw = (cw + 1, bw) if first_is_current else (cw, bw + 1)
        elif r == -1:
            cw, bw = (cw, bw + 1) if first_is_current else (cw + 1, bw)
        else:
            dr += 1
    assert [float(x) for x in out["results"]] == [float(x) for x in results]
    assert (out["current_wins"], out["best_wins"], out["draws"]) == (cw, bw, dr)
    assert out["win_ratio"] == cw / G and arena.should_promote(0.6) and not arena.should_promote(0.59)
    cur.close(); best.close()


def test_trainer_facade_trains_and_checkpoints(pkg, tmp_path):
    """AlphaZeroTrainer (src/yin_yang/ai/trainer.py) surface: train(examples, epochs, augment) -> metrics, checkpoints in the
    reference's format, and the trained weights reach the self-play evaluator."""
    import torch
    from yinyang_game_alphazero_b200 import trainer
    game = pkg["game"].YinYangGame(6, 6)
    rng = np.random.default_rng(0)
    examples = []
    for i in range(24):
        b = game.getInitBoard()
        b.board = rng.integers(-1, 2, (6, 6)).astype(np.int8)
        pi = rng.random(36); pi /= pi.sum()
        examples.append((b, pi, float(rng.choice([-1.0, 1.0]))))
    torch.manual_seed(0)
    tr = trainer.AlphaZeroTrainer(game, model_dir=str(tmp_path), batch_size=64, num_channels=128, num_res_blocks=2)
    before = {k: v.clone() for k, v in tr.nnet.state_dict().items()}
    m = tr.train(examples, epochs=6, augment=True)            # 192 samples -> 3 full batches per epoch
    assert set(m) == {"policy_loss", "value_loss", "total_loss"} and all(len(v) == 6 for v in m.values())
    assert all(np.isfinite(m["total_loss"])) and m["total_loss"][-1] < m["total_loss"][0]
    m2 = tr.train(examples[:5], epochs=1, augment=False)      # 5 samples: one remainder-sized batch
    assert np.isfinite(m2["total_loss"][0])
    after = tr.nnet.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before)
    tr.save_checkpoint(iteration=3)
    path = os.path.join(str(tmp_path), "checkpoint_3.pth.tar")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"state_dict", "board_size", "action_size"} and set(ck["state_dict"]) == set(before)
    tr2 = trainer.AlphaZeroTrainer(game, model_dir=str(tmp_path), batch_size=64, num_channels=128, num_res_blocks=2)
    tr2.load_checkpoint(iteration=3)
    for k, v in tr2.learner.state_dict().items():
        assert torch.equal(v.float(), after[k].float()), k
    p, v = tr.nnet.predict(examples[0][0])                     # the self-play evaluator sees the trained weights
    assert p.shape == (36,) and abs(p.sum() - 1) < 1e-3 and -1 <= v <= 1


@pytest.mark.parametrize("file_loop", [False, True])
def test_cli_train_mode_runs_the_whole_loop(pkg, tmp_path, file_loop):
    """train_alphazero.py --mode train (AlphaZero.run, alphazero.py:248-270): self-play -> learner -> arena -> promotion.
    Default: the examples stay on the GPU between the stages (device_loop.DeviceLoop); --file-loop: the reference's data
    files between the stages."""
    import subprocess, sys, glob
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md, dd = str(tmp_path / "models"), str(tmp_path / "data")
    r = subprocess.run([sys.executable, os.path.join(root, "train_alphazero.py"), "--mode", "train", "--rows", "6", "--cols", "6",
                        "--iterations", "1", "--episodes", "8", "--simulations", "16", "--epochs", "1", "--arena-games", "4",
                        "--model-dir", md, "--data-dir", dd] + (["--file-loop"] if file_loop else []),
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(glob.glob(os.path.join(dd, "self_play_data_*.npz"))) == (1 if file_loop else 0)
    for name in ("current_model.pth.tar", "best_model.pth.tar", "checkpoint_1.pth.tar"):
        assert os.path.exists(os.path.join(md, name)), name
    cur = torch.load(os.path.join(md, "current_model.pth.tar"), map_location="cpu", weights_only=False)["state_dict"]
    ck = torch.load(os.path.join(md, "checkpoint_1.pth.tar"), map_location="cpu", weights_only=False)["state_dict"]
    assert all(torch.equal(cur[k], ck[k]) for k in ck)                       # the trained checkpoint became the current model
    assert int(ck["bn1.num_batches_tracked"]) > 0                             # the learner actually stepped


def test_gui_ai_move_endpoint_body(pkg, tmp_path):
    """POST /api/ai_move (src/gui/server.py:30-129) without the web server: request / response dictionaries."""
    import time
    from yinyang_game_alphazero_b200 import ai_move
    game = pkg["game"].YinYangGame(6, 6)
    model = os.path.join(str(tmp_path), "best_model.pth.tar")
    pkg["network"].YinYangNeuralNetwork(game).save_model(model)
    req = {"board": [[0] * 6 for _ in range(6)], "currentPlayer": 1, "rows": 6, "cols": 6, "modelPath": model}
    out = ai_move.get_ai_move(req)
    assert out["validMove"] is True and 0 <= out["row"] < 6 and 0 <= out["col"] < 6
    board = np.zeros((6, 6), np.int8); board[0, 0] = 1; board[0, 1] = -1
    req["board"], req["currentPlayer"] = board.tolist(), -1
    t0 = time.perf_counter()
    out = ai_move.get_ai_move(req)
    dt = time.perf_counter() - t0
    b = game.getInitBoard(); b.board = board
    assert out["validMove"] is True and game.getValidMoves(b, -1)[out["row"] * 6 + out["col"]] == 1
    assert dt < 5.0
    full = np.ones((6, 6), np.int8)                                   # a pre-existing 2x2: nobody can move
    req["board"] = full.tolist()
    assert ai_move.get_ai_move(req) == {"validMove": False, "message": "No valid moves available"}
    assert "error" in ai_move.get_ai_move({"board": None})


def test_cli_train_mode_under_torchrun_two_gpus(pkg, tmp_path):
    """torchrun --nproc-per-node 2 train_alphazero.py --mode train: every rank self-plays its share, the finished games are
    all-gathered over NCCL into every rank's device replay buffer, the learner trains data parallel, rank 0 writes checkpoints,
    plays the arena and broadcasts the promoted weights (needs two GPUs; skipped on a one-GPU box)."""
    import subprocess, sys, glob, socket
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md, dd = str(tmp_path / "models"), str(tmp_path / "data")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "train_alphazero.py"), "--mode", "train", "--rows", "6", "--cols", "6",
                        "--iterations", "1", "--episodes", "16", "--simulations", "16", "--epochs", "1", "--arena-games", "4",
                        "--model-dir", md, "--data-dir", dd], capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "2 ranks" in r.stderr or "examples" in r.stderr
    cur = torch.load(os.path.join(md, "current_model.pth.tar"), map_location="cpu", weights_only=False)["state_dict"]
    ck = torch.load(os.path.join(md, "checkpoint_1.pth.tar"), map_location="cpu", weights_only=False)["state_dict"]
    assert all(torch.equal(cur[k], ck[k]) for k in ck) and int(ck["bn1.num_batches_tracked"]) > 0


def test_mcts_num_threads_is_k_leaves_per_step(pkg):
    """MCTS(num_threads=K) (mcts.py:320-321: simulations on K threads) = K leaves per game per step with virtual loss: every
    simulation is accounted for and the choice agrees with the sequential search on a clear-cut position more often than not."""
    game = pkg["game"].YinYangGame(6, 6)
    net = pkg["network"].HashStubEvaluator(game)
    seq = pkg["mcts"].MCTS(game, net, num_simulations=128, dirichlet_noise=False, verbose=0)
    par = pkg["mcts"].MCTS(game, net, num_simulations=128, num_threads=8, dirichlet_noise=False, verbose=0)
    rng = np.random.default_rng(3)
    boards = np.zeros((6, 6, 6), np.int8)
    for i in range(6):
        boards[i, rng.integers(0, 6), rng.integers(0, 6)] = 1
    players = -np.ones(6, np.int8)
    c1, _ = seq.search_batch(boards, players)
    c8, _ = par.search_batch(boards, players)
    assert np.array_equal(c8.sum(axis=1), c1.sum(axis=1)) and np.all(c8.sum(axis=1) == 128)
    top3 = [len(set(np.argsort(-c1[i])[:3]) & set(np.argsort(-c8[i])[:3])) for i in range(6)]
    assert np.mean(top3) >= 1.5
    seq.close(); par.close()


def test_mcts_facade_reference_unit_cases(pkg, oracle_mod):
    """The reference's own MCTS unit cases (src/yin_yang/ai/mcts_tests.py) that are about the search rather than about
    its fake game, on real boards through the facade: root visits == simulations with and without Dirichlet noise
    (:215-235), children carry the evaluator's raw priors (mcts.py:77-78), a node's visits are its own first visit plus
    its children's (:389-416, backpropagation), sharper distributions at lower temperatures (:418-445), a single legal
    move gets probability 1 (:477-496)."""
    from conftest import random_play_boards
    game = pkg["game"].YinYangGame(6, 6)
    stub = pkg["network"].HashStubEvaluator(game)
    mcts = pkg["mcts"].MCTS(game, stub, num_simulations=60, verbose=0)
    board = game.getInitBoard()
    board.board[2, 2], board.board[2, 3] = 1, -1
    probs, root = mcts.search(board, 1)
    assert abs(probs.sum() - 1.0) < 1e-12 and root.is_expanded() and root.visits == 60
    pol, _ = stub.predict(board)
    legal = game.getValidMoves(board, 1)
    assert sorted(root.children) == list(np.flatnonzero(legal))                     # one child per legal action, ascending
    for a, ch in root.children.items():
        assert ch.prior == pol[a] and isinstance(ch.value_sum, np.float32)
        assert ch.parent is root and ch.action == a
        if ch.visits > 0:
            assert ch.player == -1 and np.abs(ch.board.board).sum() == 3          # expanded: holds its own position
            if not ch.is_terminal and ch.children:
                assert ch.visits == 1 + sum(g.visits for g in ch.children.values())
        else:
            assert not ch.is_expanded() and ch.board is None
    assert sum(ch.visits for ch in root.children.values()) == 60
    deep = max(root.children.values(), key=lambda c: c.visits)
    assert any(g.visits > 0 and g.is_expanded() for g in deep.children.values())   # the tree is reachable below depth 1
    probs_n, root_n = mcts.search(board, 1, add_exploration_noise=True)
    assert abs(probs_n.sum() - 1.0) < 1e-12 and root_n.visits == 60
    d1, d2, d3, d4 = (root.get_children_distribution(t) for t in (1.0, 0.5, 0.1, 0))
    order = np.argsort(-root.get_children_visit_counts())
    a0, a1 = order[0], order[1]
    if root.get_children_visit_counts()[a0] > root.get_children_visit_counts()[a1]:
        assert d1[a0] - d1[a1] < d2[a0] - d2[a1] < d3[a0] - d3[a1] and d4[a0] == 1.0
    sharp = pkg["mcts"].MCTS(game, stub, num_simulations=60, temperature=0.5, verbose=0)
    p_sharp, _ = sharp.search(board, 1)
    assert np.allclose(p_sharp, d2)                                                 # search returns the distribution at MCTS.temperature (mcts.py:323)
    # a position with exactly one legal move for the side to move
    boards, players = random_play_boards(oracle_mod, 6, 6, 3000, seed=4)
    masks = oracle_mod.legal_mask(boards, players, 6, 6)
    one = np.flatnonzero(masks.sum(axis=1) == 1)
    assert one.size > 0
    b1 = game.getInitBoard(); b1.board = boards[one[0]].copy()
    p1, r1 = mcts.search(b1, int(players[one[0]]))
    only = int(np.flatnonzero(masks[one[0]])[0])
    assert np.argmax(p1) == only and abs(p1[only] - 1.0) < 1e-12 and list(r1.children) == [only]
    mcts.close(); sharp.close()


def test_mcts_facade_sees_reloaded_weights(pkg):
    """The reference MCTS holds the network by reference (mcts.py:249): a search after neural_net.load_state_dict must use
    the new weights, although the facade's engine keeps its own packed copy on the device."""
    import torch
    game = pkg["game"].YinYangGame(4, 4)
    torch.manual_seed(1)
    net = pkg["network"].YinYangNeuralNetwork(game, num_channels=32, num_res_blocks=1)
    mcts = pkg["mcts"].MCTS(game, net, num_simulations=40, dirichlet_noise=False, verbose=0)
    board = game.getInitBoard()
    c0 = mcts.search(board, 1)[1].get_children_visit_counts()
    torch.manual_seed(2)
    other = pkg["network"].YinYangNeuralNetwork(game, num_channels=32, num_res_blocks=1)
    fresh = pkg["mcts"].MCTS(game, other, num_simulations=40, dirichlet_noise=False, verbose=0)
    c_other = fresh.search(board, 1)[1].get_children_visit_counts()
    net.load_state_dict(other.state_dict())
    c1 = mcts.search(board, 1)[1].get_children_visit_counts()
    assert np.array_equal(c1, c_other) and not np.array_equal(c0, c_other)
    mcts.close(); fresh.close()


def test_self_play_array_wire_and_rolling_driver(pkg, tmp_path):
    """wire='arrays' (boards as one int8 array under the reference's npz keys) loads in this package's TrainingDataQueue;
    RollingSelfPlay returns whole finished games, keeps the other slots' games in flight across calls (keep_warm)."""
    import torch
    from yinyang_game_alphazero_b200.training_pipeline import TrainingDataQueue
    game = pkg["game"].YinYangGame(4, 4)
    torch.manual_seed(0)
    net = pkg["network"].YinYangNeuralNetwork(game, num_channels=32, num_res_blocks=1)
    model = str(tmp_path / "m" / "best_model.pth.tar")
    net.save_model(model)
    path = pkg["self_play"].generate_self_play_data(game, model, str(tmp_path / "d"), num_games=5, num_simulations=12, wire="arrays")
    d = np.load(path)                                                              # no pickle needed
    assert d["boards"].dtype == np.int8 and d["boards"].shape[1:] == (4, 4) and len(d["values"]) == len(d["boards"]) >= 5
    q = TrainingDataQueue()
    q.push_file(path)
    assert len(q) == len(d["values"])
    sp = pkg["self_play"].RollingSelfPlay(game, net.state_dict(), slots=8, num_simulations=12, seed=3)
    first = sp.play(6, keep_warm=True)
    second = sp.play(6, keep_warm=True)
    for ex in (first, second):
        games = np.unique(ex["game"])
        assert len(games) >= 6
        for g in games:
            sel = ex["game"] == g
            assert np.array_equal(ex["ply"][sel], np.arange(sel.sum())) and len(set(ex["z"][sel])) == 1   # whole games, one z
            assert np.abs(ex["boards"][sel][0]).sum() == 0
    assert not set(np.unique(first["game"])) & set(np.unique(second["game"]))
    sp.close()


def test_cli_self_play_mode_under_torchrun_two_gpus(pkg, tmp_path):
    """torchrun --nproc-per-node 2 train_alphazero.py --mode self-play: the episodes are sharded over the ranks, the examples
    gathered to rank 0 over NCCL, ONE data file is written (needs two GPUs)."""
    import subprocess, sys, glob, socket
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md, dd = str(tmp_path / "models"), str(tmp_path / "data")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "train_alphazero.py"), "--mode", "self-play", "--rows", "6", "--cols", "6",
                        "--episodes", "21", "--simulations", "16", "--init-model", "--model-dir", md, "--data-dir", dd],
                       capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    files = glob.glob(os.path.join(dd, "self_play_data_*.npz"))
    assert len(files) == 1
    d = np.load(files[0])
    assert d["boards"].shape[1:] == (6, 6) and len(d["values"]) >= 21
    assert int((np.abs(d["boards"]).sum(axis=(1, 2)) == 0).sum()) == 21          # one empty first position per game: 11 + 10 games


def test_device_loop_single_gpu(pkg):
    """DeviceLoop on one GPU: two iterations, examples go replay ring -> device buffer -> augmentation -> learner without
    leaving the device; rolling self-play returns at least the requested games; the losses are finite and the weights move."""
    import torch
    from yinyang_game_alphazero_b200.device_loop import DeviceLoop
    game = pkg["game"].YinYangGame(4, 4)
    torch.manual_seed(0)
    loop = DeviceLoop(game, num_iterations=2, num_episodes=12, num_simulations=16, num_epochs=1, batch_size=32, sample_size=256,
                      eval_games=4, num_channels=32, num_res_blocks=1, slots=8, save_files=False, seed=1)
    w0 = loop.learner.params.clone()
    hist = loop.run()
    assert len(hist) == 2 and all(h["examples_global"] > 0 and h["learner_steps"] > 0 for h in hist)
    assert all(np.isfinite(h["policy_loss"]) and np.isfinite(h["value_loss"]) for h in hist)
    assert len(loop.buffer) == sum(h["examples_global"] for h in hist)
    assert not torch.equal(w0, loop.learner.params)
    assert set(torch.unique(loop.buffer.z[: len(loop.buffer)]).tolist()) <= {1.0, -1.0, float(np.float32(0.0001))}
    loop.close()
