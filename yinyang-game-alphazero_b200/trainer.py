"""Host-side mirror of src/yin_yang/ai/trainer.py (AlphaZeroTrainer) on the GPU learner.

Same constructor arguments, ``train(examples, epochs, augment)`` with the reference's metrics dictionary, and
``save_checkpoint`` / ``load_checkpoint`` writing / reading the reference's checkpoint files (trainer.py:163-207,
neural_network.py:198-237).  The dataset is built on the device (data_utils.dataset_tensors -> yy_augment_samples), each
epoch is one ``torch.randperm`` shuffle (DataLoader(shuffle=True), trainer.py:96-100) cut into batches of ``batch_size``
with a smaller last batch, and every batch is one ``Learner.step`` (a CUDA-graph replay of hand-written kernels).
"""
from __future__ import annotations

import logging
import os
import time

import torch

from . import data_utils
from .learner import Learner
from .network import YinYangNeuralNetwork

logger = logging.getLogger("YinYangNN.trainer")


class AlphaZeroTrainer:
    def __init__(self, game, model_dir="models", lr=0.001, batch_size=64, weight_decay=1e-4, device=None,
                 num_channels=128, num_res_blocks=10, precision="3xtf32", data_parallel=True):
        self.game = game
        self.model_dir = model_dir
        self.batch_size = batch_size
        if not os.path.exists(model_dir):
            os.makedirs(model_dir)
        self.device = torch.device(device if device is not None else "cuda")
        self.nnet = YinYangNeuralNetwork(game, num_channels, num_res_blocks)        # reference initialisation (trainer.py:48)
        n, m = game.getBoardSize()
        self.learner = Learner(n, m, num_channels, num_res_blocks, batch_size=batch_size, lr=lr, weight_decay=weight_decay,
                               state_dict=self.nnet.state_dict(), device=self.device, precision=precision, data_parallel=data_parallel)
        logger.info(f"Training parameters: lr={lr}, batch_size={batch_size}, weight_decay={weight_decay}")

    def train(self, examples, epochs=10, augment=True):
        """trainer.py:67-161.  examples: list of (board, policy, value).  Returns {'policy_loss': [...], 'value_loss': [...],
        'total_loss': [...]} with one sample-weighted mean per epoch."""
        logger.info(f"Training on {len(examples)} examples for {epochs} epochs")
        planes, policies, values = data_utils.dataset_tensors(examples, self.game, augment=augment)
        n = planes.shape[0]
        if self.learner.world > 1:
            # data parallel: every step is a collective (one all-reduce of the gradients), so all ranks must run the same
            # number of steps per epoch -- agree on the smallest dataset and drop the surplus samples of the others
            import torch.distributed as dist
            t = torch.tensor([n], dtype=torch.int64, device=planes.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            n = int(t.item())
            planes, policies, values = planes[:n], policies[:n], values[:n]
        metrics = {"policy_loss": [], "value_loss": [], "total_loss": []}
        for epoch in range(epochs):
            t0 = time.time()
            perm = torch.randperm(n, device=planes.device)
            pl, po, va = planes[perm], policies[perm], values[perm]    # one gather per epoch; batches are contiguous slices
            acc = torch.zeros(2, dtype=torch.float64, device=planes.device)
            for i in range(0, n, self.batch_size):
                j = min(n, i + self.batch_size)
                losses = self.learner.step(pl[i:j], po[i:j], va[i:j])
                acc += losses.double() * (j - i)                       # stays on the device: no sync per batch
            pl, vl = (acc / max(n, 1)).tolist()
            metrics["policy_loss"].append(pl); metrics["value_loss"].append(vl); metrics["total_loss"].append(pl + vl)
            logger.info(f"Epoch {epoch + 1}/{epochs} - Policy Loss: {pl:.4f}, Value Loss: {vl:.4f}, "
                        f"Total Loss: {pl + vl:.4f}, Time: {time.time() - t0:.2f}s")
        self.nnet.load_state_dict(self.learner.state_dict())
        return metrics

    def _path(self, filename, iteration):
        if filename is None:
            filename = f"checkpoint_{iteration}.pth.tar" if iteration is not None else "checkpoint.pth.tar"
        return os.path.join(self.model_dir, filename)

    def save_checkpoint(self, filename=None, iteration=None):  # trainer.py:163-179
        self.nnet.load_state_dict(self.learner.state_dict())
        self.nnet.save_model(self._path(filename, iteration))

    def load_checkpoint(self, filename=None, iteration=None):  # trainer.py:181-207
        path = self._path(filename, iteration)
        if os.path.exists(path):
            self.nnet.load_model(path)
            self.learner.load_state_dict(self.nnet.state_dict())
            logger.info(f"Loaded model from {path}")
        else:
            logger.warning(f"No checkpoint found at {path}")
