"""Developer tool: two eager learner steps (no CUDA graph) for an ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import learner as lrn
from yinyang_game_alphazero_b200 import network
torch.manual_seed(0)
prec = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
net = network._Params(8, 8, 128, 10)
L = lrn.Learner(8, 8, 128, 10, batch_size=64, state_dict=net.state_dict(), use_graph=False, precision=prec)
planes = torch.rand(64, 5, 8, 8, device="cuda"); pi = torch.softmax(torch.randn(64, 64, device="cuda"), 1); z = torch.rand(64, device="cuda") * 2 - 1
for _ in range(2):
    L.step(planes, pi, z)
torch.cuda.synchronize()
