"""Developer diagnostic (test infrastructure: it runs the oracle as the checker, so it lives under tests/): kernel vs numpy
emulation vs fp32 torch, error statistics per configuration.  Run by hand on a GPU box: `python tests/nn_diag.py`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine, weights
from oracle import port
import oracle
import emulate_tower as emu
from conftest import random_play_boards, randomise_bn

for (n, m, C, blocks, count, rnd) in [(8, 8, 128, 0, 24, True), (8, 8, 128, 1, 24, True), (8, 8, 128, 2, 24, True), (8, 8, 128, 10, 24, True),
                                      (8, 8, 128, 10, 24, False), (6, 6, 128, 3, 24, True), (16, 16, 128, 1, 9, True), (5, 7, 64, 2, 24, True)]:
    torch.manual_seed(0)
    net = port.build_net(n, m, C, blocks)
    net = randomise_bn(net) if rnd else net.eval()
    e = engine.Engine(rows=n, cols=m, n_games=max(count, 4), n_sims=1, evaluator="nn", state_dict=net.state_dict())
    boards, _ = random_play_boards(oracle, n, m, count, seed=9)
    policy, value, logits = e.evaluate_host(boards, want_logits=True)
    with torch.no_grad():
        rl, rv = net(net.planes(boards))
    rl, rv = rl.numpy(), rv.numpy()[:, 0]
    img = weights.pack_state_dict(net.state_dict(), n, m); lay = weights.layout(n, m, C, blocks)
    el, ev, _ = emu.forward(img, lay, n, m, blocks, boards)
    f = lambda a, b: (float(np.abs(a - b).max()), float(np.abs(a - b).mean()))
    print((n, m, C, blocks, rnd), "scale", float(np.abs(rl).max()), "std", float(rl.std()),
          "| kern-emu", f(logits, el), f(value, ev), "| kern-fp32", f(logits, rl), f(value, rv), "| emu-fp32", f(el, rl), f(ev, rv),
          "| top1", float((logits.argmax(1) == rl.argmax(1)).mean()), flush=True)
    e.close()
