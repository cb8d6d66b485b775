"""Developer tool: the same training step with stock PyTorch on the same GPU (cuDNN convolutions, torch autograd, torch.optim.Adam;
eager, TF32 convolutions on/off) -- a library baseline next to the hand-written learner step."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import network

torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = network._Params(8, 8, 128, 10).cuda().train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
planes = torch.rand(B, 5, 8, 8, device="cuda"); pi = torch.softmax(torch.randn(B, 64, device="cuda"), 1); z = torch.rand(B, device="cuda") * 2 - 1
ce, mse = torch.nn.CrossEntropyLoss(), torch.nn.MSELoss()


def step():
    opt.zero_grad(set_to_none=True)
    lg, v = net(planes)
    loss = ce(lg, pi) + mse(v.view(-1), z)
    loss.backward()
    opt.step()


out = {}
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(50):
        step()
    ev1.record(); torch.cuda.synchronize()
    out["cudnn_tf32" if tf32 else "cudnn_fp32"] = {"ms_per_step": ev0.elapsed_time(ev1) / 50, "samples_per_s": B * 50 / (ev0.elapsed_time(ev1) * 1e-3)}
print(json.dumps({"batch": B, "torch": torch.__version__, **out}))
