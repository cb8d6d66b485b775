"""Pins the oracle (C restatement + Python port) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_files, load_golden


@pytest.mark.parametrize("name", golden_files("rules_"))
def test_c_oracle_rules_match_reference(oracle_mod, name):
    g = load_golden(name)
    n, m, B = int(g["n"]), int(g["m"]), g["boards"]
    N = len(B)
    one = np.ones(N, np.int8)
    assert np.array_equal(oracle_mod.legal_mask(B, one, n, m), g["mask_black"])
    assert np.array_equal(oracle_mod.legal_mask(B, -one, n, m), g["mask_white"])
    nb, npl = oracle_mod.next_state(B, g["players"], g["actions"], n, m)
    assert np.array_equal(nb, g["next_boards"]) and np.array_equal(npl, g["next_players"])
    assert np.array_equal(oracle_mod.game_ended(B, one, n, m), g["ended_black"])
    assert np.array_equal(oracle_mod.game_ended(B, -one, n, m), g["ended_white"])


@pytest.mark.parametrize("name", golden_files("mcts_"))
def test_c_oracle_mcts_matches_reference(oracle_mod, name):
    g = load_golden(name)
    noise = g["noise"] if g["noise"].size else None
    r = oracle_mod.mcts_search(g["board"], int(g["player"]), int(g["n"]), int(g["m"]), int(g["sims"]),
                               cpuct=float(g["cpuct"]), noise=noise)
    assert np.array_equal(r["counts"], g["counts"])
    assert np.array_equal(r["child_w"], g["child_w"])          # float32 value sums, bit for bit
    assert r["n_evals"] == int(g["n_evals"])


def test_python_port_rules_match_reference():
    from oracle import port
    g = load_golden("rules_6x6.npz")
    game = port.Game(6, 6)
    for i in range(0, len(g["boards"]), 7):
        b = port.Board(6, 6, g["boards"][i])
        assert np.array_equal(game.getValidMoves(b, 1).astype(np.uint8), g["mask_black"][i])
        assert np.array_equal(game.getValidMoves(b, -1).astype(np.uint8), g["mask_white"][i])
        nb, npl = game.getNextState(b, int(g["players"][i]), int(g["actions"][i]))
        assert np.array_equal(nb.board, g["next_boards"][i]) and npl == g["next_players"][i]
        assert game.getGameEnded(b, 1) == g["ended_black"][i]


@pytest.mark.parametrize("name", ["mcts_4x4_s300.npz", "mcts_6x6_s100_noise.npz", "mcts_4x4_s300_mid.npz"])
def test_python_port_mcts_matches_reference(oracle_mod, name):
    from oracle import port
    g = load_golden(name)
    n, m = int(g["n"]), int(g["m"])
    noise = g["noise"] if g["noise"].size else None
    counts, root = port.search(port.Game(n, m), port.HashStubNet(n, m), port.Board(n, m, g["board"]), int(g["player"]),
                               int(g["sims"]), cpuct=float(g["cpuct"]), noise=noise)
    assert np.array_equal(counts.astype(np.int32), g["counts"])


@pytest.mark.parametrize("name", golden_files("net_"))
def test_python_port_network_matches_reference(name):
    import torch
    from oracle import port
    g = load_golden(name)
    n, m = int(g["n"]), int(g["m"])
    net = port.build_net(n, m, int(g["channels"]), int(g["blocks"]))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    net.load_state_dict(sd)
    net.eval()
    x = net.planes(g["boards"])
    assert np.array_equal(x.numpy(), g["planes"])               # board_to_input, exact
    with torch.no_grad():
        lg, v = net(x)
    np.testing.assert_allclose(lg.numpy(), g["logits"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(v.numpy()[:, 0], g["values"], rtol=0, atol=2e-6)
    p, pv = net.predict(port.Board(n, m, g["boards"][3]))
    np.testing.assert_allclose(p, g["policies"][3], rtol=0, atol=1e-6)


def test_stub_evaluator_spec(oracle_mod):
    """The hash stub used for the MCTS goldens (pure Python ints in make_golden.py) == the C one."""
    g = load_golden("mcts_6x6_s400.npz")
    pol, val = oracle_mod.stub_predict(g["board"], 6, 6)
    assert pol.dtype == np.float32 and np.all(pol > 0) and np.all(pol <= 4096 / 65536)
    assert np.all(pol * 65536 == np.round(pol * 65536))        # dyadic
    assert -1.0 <= float(val) < 1.0 and float(val) * 65536 == round(float(val) * 65536)


@pytest.mark.parametrize("name", golden_files("augment_"))
def test_oracle_augmentation_matches_reference(oracle_mod, name):
    """create_dataset_from_games / DataProcessor.augment_sample (data_utils.py:39-215): bit-exact float32 planes,
    policies and values for all 8 symmetric forms."""
    g = load_golden(name)
    n, m = int(g["n"]), int(g["m"])
    planes, policies, values = oracle_mod.augment_dataset(g["boards"], g["counts"], g["values"], n, m)
    assert np.array_equal(planes, g["planes"])
    assert np.array_equal(policies, g["policies"])
    assert np.array_equal(values, g["out_values"])


@pytest.mark.parametrize("name", golden_files("selfplay_"))
def test_oracle_play_game_matches_reference_golden(oracle_mod, name):
    """The oracle's restatement of SelfPlayWorker.play_game (self_play.py:72-192) against whole games played by the
    unmodified reference on the same recorded random stream: every stored position, pi (float64) and z."""
    g = load_golden(name)
    n, m = int(g["n"]), int(g["m"])
    noise = g["noise"] if g["noise"].size and float(g["eps"]) > 0 else None
    b, pi, z, kind = oracle_mod.play_game(n, m, int(g["sims"]), g["uniforms"], noise, cpuct=float(g["cpuct"]), eps=float(g["eps"]),
                                          temperature_threshold=int(g["temperature_threshold"]))
    assert np.array_equal(b, g["boards"]) and np.array_equal(pi, g["pis"]) and np.array_equal(z, g["zs"])
    assert kind in ("ended", "passes")


@pytest.mark.parametrize("name", golden_files("rowcol_"))
def test_oracle_rowcol_rule_matches_js_transcription(oracle_mod, name):
    """YY_RULE_ROWCOL against tests/golden/rowcol_*.npz: the browser game's isValidMove (yin_yang_game.js:186-232) with
    checkRowColumnConstraint (:338-384) transcribed loop for loop on top of the unmodified Python reference's
    is_valid_move (tests/golden/make_golden_rowcol.py)."""
    g = load_golden(name)
    n, m, B = int(g["n"]), int(g["m"]), g["boards"]
    one = np.ones(len(B), np.int8)
    assert np.array_equal(oracle_mod.legal_mask(B, one, n, m, 1), g["mask_black"])
    assert np.array_equal(oracle_mod.legal_mask(B, -one, n, m, 1), g["mask_white"])


def test_rowcol_rule_hand_worked_boards(oracle_mod):
    """Boards worked out by hand from yin_yang_game.js:338-384 ("a row / column with no empty cell and only one colour is
    a violation", checked on the whole board after the trial placement, :220-227).  X black, O white, . empty, 4x4:
      A  XXX.   black at (0,3) would fill row 0 with black only          -> banned with the flag, legal without
         O...   (connectivity: (0,3) touches (0,2); 2x2: no window)
      B  XXO.   black at (0,3): row 0 = X X O X, two colours              -> legal either way
      C  X...   black at (3,0) would fill column 0 with black only        -> banned with the flag
         X...
         X...
         ..O.
      D  OOOO   row 0 is ALREADY all white: the check runs over the whole board after any placement, so with the flag
         X...   nobody has a legal move anywhere (:338-358 returns false whatever was placed)
      E  XXX.   white at (0,3): row 0 = X X X O, two colours -> legal for white (it touches white's only stone (1,3))
         ...O
    """
    def board(rows):
        return np.array([[{"X": 1, "O": -1, ".": 0}[c] for c in r] for r in rows], np.int8)[None]
    one = np.ones(1, np.int8)
    A = board(["XXX.", "O...", "....", "...."])
    assert oracle_mod.legal_mask(A, one, 4, 4, 0)[0, 3] == 1 and oracle_mod.legal_mask(A, one, 4, 4, 1)[0, 3] == 0
    Bb = board(["XXO.", "....", "....", "...."])
    assert oracle_mod.legal_mask(Bb, one, 4, 4, 0)[0, 3] == oracle_mod.legal_mask(Bb, one, 4, 4, 1)[0, 3] == 0   # (0,3) does not touch black
    Bb = board(["XXO.", "...X", "....", "...."])                                   # two black groups: (0,3) joins nothing to (0,0)
    assert oracle_mod.legal_mask(Bb, one, 4, 4, 1)[0, 3] == 0
    B2 = board(["XOX.", "XXX.", "....", "...."])                                   # black at (0,3): row X O X X, touches (1,3)? no: (0,2) yes
    assert oracle_mod.legal_mask(B2, one, 4, 4, 0)[0, 3] == 1 and oracle_mod.legal_mask(B2, one, 4, 4, 1)[0, 3] == 1
    C = board(["X...", "X...", "X...", "..O."])
    assert oracle_mod.legal_mask(C, one, 4, 4, 0)[0, 12] == 1 and oracle_mod.legal_mask(C, one, 4, 4, 1)[0, 12] == 0
    assert oracle_mod.legal_mask(C, one, 4, 4, 1)[0, 1] == 1                       # (0,1): row 0 = X X . . , not full
    D = board(["OOOO", "X...", "....", "...."])
    assert oracle_mod.legal_mask(D, one, 4, 4, 0)[0].sum() > 0
    assert oracle_mod.legal_mask(D, one, 4, 4, 1)[0].sum() == 0 and oracle_mod.legal_mask(D, -one, 4, 4, 1)[0].sum() == 0
    E = board(["XXX.", "...O", "....", "...."])
    assert oracle_mod.legal_mask(E, -one, 4, 4, 1)[0, 3] == 1 and oracle_mod.legal_mask(E, one, 4, 4, 1)[0, 3] == 0
