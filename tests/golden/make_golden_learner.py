#!/usr/bin/env python
"""Golden fixtures for the learner step from the UNMODIFIED Python reference.

Run in the build container only (it imports /root/reference):

    cd /tmp && PYTHONDONTWRITEBYTECODE=1 python /root/repo/tests/golden/make_golden_learner.py

learner_<n>x<m>_c<C>_b<blocks>.npz pins three optimisation steps of the reference's own ``AlphaZeroTrainer`` -- its
network (``trainer.nnet``, in ``train()`` mode), its loss objects (``policy_loss_fn`` / ``value_loss_fn``) and its Adam
optimizer (``trainer.optimizer``), called exactly as the inner loop of ``AlphaZeroTrainer.train`` calls them
(src/yin_yang/ai/trainer.py:120-137) -- on fixed batches (a full batch twice, then a smaller remainder batch):
  init.*        state_dict before the first step (random weights, batch-norm affine parameters moved off their defaults)
  planes/pi/z   the batches (planes from the reference's board_to_input)
  loss_p/loss_v the losses of each step
  grad0.*       the gradients of the first step (p.grad after total_loss.backward()); gnorm = per-tensor gradient norms of
                every step
  after.*       the state_dict after the third optimizer.step()
"""
import logging
import os
import sys

import numpy as np

REF = os.environ.get("YY_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
logging.disable(logging.CRITICAL)

import torch  # noqa: E402
from src.yin_yang import YinYangGame  # noqa: E402
from src.yin_yang.ai.trainer import AlphaZeroTrainer  # noqa: E402
from src.yin_yang.ai.neural_network import YinYangNeuralNetwork  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def make(n, m, channels, blocks, batch, seed):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    game = YinYangGame(n, m)
    tr = AlphaZeroTrainer(game, model_dir="/tmp/yy_golden_models", lr=0.001, batch_size=batch, weight_decay=1e-4, device=torch.device("cpu"))
    tr.nnet = YinYangNeuralNetwork(game, num_channels=channels, num_res_blocks=blocks)       # the trainer's network, smaller
    tr.optimizer = torch.optim.Adam(tr.nnet.parameters(), lr=0.001, weight_decay=1e-4)       # trainer.py:52-56
    with torch.no_grad():
        for k, p in tr.nnet.named_parameters():
            if "bn" in k or k.endswith("bias"):
                p.add_(torch.randn_like(p) * 0.2)
    out = {f"init.{k}": v.numpy().copy() for k, v in tr.nnet.state_dict().items()}
    boards = []
    for _ in range(batch):
        b = game.getInitBoard()
        b.board[:, :] = rng.integers(-1, 2, (n, m)).astype(np.int8)
        boards.append(b)
    planes = torch.stack([tr.nnet.board_to_input(b) for b in boards])                       # neural_network.py:156-196
    pi = rng.random((batch, n * m)); pi /= pi.sum(1, keepdims=True)
    pi = torch.tensor(pi, dtype=torch.float32)
    z = torch.tensor(rng.choice([-1.0, 1.0, 0.0001], batch), dtype=torch.float32)
    out.update(planes=planes.numpy(), pi=pi.numpy(), z=z.numpy(), grids=np.stack([b.board for b in boards]))
    sizes = [batch, batch, batch - 3]
    lp, lv, gnorm = [], [], []
    tr.nnet.train()                                                                           # trainer.py:110
    for k, b in enumerate(sizes):
        bo, po, va = planes[:b], pi[:b], z[:b]
        tr.optimizer.zero_grad()                                                              # trainer.py:126-137
        policy_logits, value_preds = tr.nnet(bo)
        policy_loss = tr.policy_loss_fn(policy_logits, po)
        value_loss = tr.value_loss_fn(value_preds.view(-1), va)
        total_loss = policy_loss + value_loss
        total_loss.backward()
        if k == 0:
            for name, p in tr.nnet.named_parameters():
                out[f"grad0.{name}"] = p.grad.numpy().copy()
        gnorm.append([p.grad.norm().item() for _, p in tr.nnet.named_parameters()])
        tr.optimizer.step()
        lp.append(policy_loss.item()); lv.append(value_loss.item())
    for name, v in tr.nnet.state_dict().items():
        out[f"after.{name}"] = v.numpy().copy()
    out.update(loss_p=np.array(lp), loss_v=np.array(lv), sizes=np.array(sizes), channels=channels, blocks=blocks, gnorm=np.array(gnorm),
               names=np.array([k for k, _ in tr.nnet.named_parameters()]))
    path = os.path.join(OUT, f"learner_{n}x{m}_c{channels}_b{blocks}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KB", lp, lv)


if __name__ == "__main__":
    make(4, 4, 8, 2, 8, seed=11)
