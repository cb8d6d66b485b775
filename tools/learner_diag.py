"""Developer tool: CUDA learner step vs the same layer sequence on the torch emulation of the kernels (CPU fp32):
per-buffer and per-gradient relative differences, to localise a numerical problem."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import learner
from emu_learner_ops import TorchEmuOps
from test_learner import _reference_net, _batch

n, C, nb, B = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (6, 32, 2, 20)))
prec = sys.argv[5] if len(sys.argv) > 5 else '3xtf32'
net = _reference_net(n, n, C, nb, seed=7)
planes, pi, z = _batch(net, n, n, B, seed=8)
Lc = learner.Learner(n, n, C, nb, batch_size=B, state_dict=net.state_dict(), use_graph=False, precision=prec)
Le = learner.Learner(n, n, C, nb, batch_size=B, state_dict=net.state_dict(), _ops=TorchEmuOps())
Lc.step(planes.cuda(), pi.cuda(), z.cuda()); Le.step(planes, pi, z)
rel = lambda a, b: ((a.cpu() - b).abs().max() / (b.abs().max() + 1e-20)).item()
for i in range(1 + 2 * nb):
    print(f"Y[{i}] {rel(Lc.Y[i], Le.Y[i]):.2e}  act[{i}] {rel(Lc.act[i], Le.act[i]):.2e}")
for h in ("policy", "value"):
    print(h, f"Yh {rel(Lc.Yh[h], Le.Yh[h]):.2e} acth {rel(Lc.acth[h], Le.acth[h]):.2e}")
print(f"logits {rel(Lc.logits, Le.logits):.2e} hid {rel(Lc.hid, Le.hid):.2e} dlogits {rel(Lc.dlogits, Le.dlogits):.2e} dhid {rel(Lc.dhid, Le.dhid):.2e}")
gc, ge = Lc.grad_dict(), Le.grad_dict()
l2 = lambda a, b: ((a.cpu() - b).norm() / (b.norm() + 1e-20)).item()
cos = lambda a, b: torch.nn.functional.cosine_similarity(a.cpu().flatten(), b.flatten(), dim=0).item()
print("relu-mask flips: hid", int(((Lc.hid.cpu() > 0) != (Le.hid > 0)).sum()), "of", Le.hid.numel(),
      "; act:", [int(((Lc.act[i].cpu() > 0) != (Le.act[i] > 0)).sum()) for i in range(1 + 2 * nb)], "of", Le.act[0].numel())
for k in ge:
    print(f"grad {k:34s} max-rel {rel(gc[k], ge[k]):.2e}  l2-rel {l2(gc[k], ge[k]):.2e}  cos {cos(gc[k], ge[k]):.6f}  max|g| {ge[k].abs().max().item():.2e}")
allc = torch.cat([gc[k].flatten() for k in ge if not (k.endswith('bias') and 'conv' in k)]); alle = torch.cat([ge[k].flatten() for k in ge if not (k.endswith('bias') and 'conv' in k)])
print("whole gradient: l2-rel", l2(allc, alle), "cos", cos(allc, alle))
