"""Host-side mirror of src/yin_yang/ai/training_pipeline.py: TrainingDataQueue, TrainingPipeline, run_training_pipeline.

Same classes, arguments and file conventions (``self_play_data_*.npz`` in, ``checkpoint_<iteration>.pth.tar`` out); the
trainer underneath is the GPU learner (trainer.AlphaZeroTrainer -> learner.Learner).
"""
from __future__ import annotations

import glob
import logging
import os
import random
from collections import deque

import numpy as np

from .trainer import AlphaZeroTrainer

logger = logging.getLogger("YinYangTraining")


def _rank():
    try:
        import torch.distributed as dist
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    except Exception:
        return 0


class TrainingDataQueue:  # training_pipeline.py:23-106
    def __init__(self, max_size=500000, sample_size=10000):
        self.max_size = max_size
        self.sample_size = min(sample_size, max_size)
        self.queue = deque(maxlen=max_size)

    def push_examples(self, examples):
        old = len(self.queue)
        for ex in examples:
            self.queue.append(ex)
        logger.info(f"Added {len(self.queue) - old} examples to queue. Queue size: {len(self.queue)}")

    def push_file(self, file_path):
        if not os.path.exists(file_path):
            logger.error(f"File {file_path} not found")
            return
        data = np.load(file_path, allow_pickle=True)
        boards, policies, values = data["boards"], data["policies"], data["values"]
        self.push_examples([(boards[i], policies[i], values[i]) for i in range(len(boards))])
        logger.info(f"Loaded {len(boards)} examples from {file_path}")

    def sample(self, sample_size=None):
        if sample_size is None:
            sample_size = self.sample_size
        if len(self.queue) == 0:
            return []
        return random.sample(list(self.queue), min(sample_size, len(self.queue)))

    def __len__(self):
        return len(self.queue)


class TrainingPipeline:  # training_pipeline.py:108-287
    def __init__(self, game, model_dir="models", data_dir="data", lr=0.001, batch_size=64, weight_decay=1e-4,
                 epochs_per_iteration=10, sample_size=10000, queue_size=500000, checkpoint_interval=10, **trainer_kwargs):
        self.game, self.model_dir, self.data_dir = game, model_dir, data_dir
        self.epochs_per_iteration, self.sample_size, self.checkpoint_interval = epochs_per_iteration, sample_size, checkpoint_interval
        for d in (model_dir, data_dir):
            os.makedirs(d, exist_ok=True)
        self.trainer = AlphaZeroTrainer(game=game, model_dir=model_dir, lr=lr, batch_size=batch_size, weight_decay=weight_decay,
                                        **trainer_kwargs)
        self.data_queue = TrainingDataQueue(max_size=queue_size, sample_size=sample_size)
        self.iteration = 0
        self._load_iteration()

    def _load_iteration(self):
        cps = glob.glob(os.path.join(self.model_dir, "checkpoint_*.pth.tar"))
        if not cps:
            return
        latest = max(int(os.path.basename(cp).split("_")[1].split(".")[0]) for cp in cps)
        self.trainer.load_checkpoint(iteration=latest)
        self.iteration = latest

    def load_data(self):
        files = sorted(glob.glob(os.path.join(self.data_dir, "self_play_data_*.npz")))
        if not files:
            logger.warning("No data files found.")
            return
        for f in files:
            self.data_queue.push_file(f)

    def train_iteration(self):
        examples = self.data_queue.sample()
        if not examples:
            logger.warning("No examples available for training.")
            return {}
        metrics = self.trainer.train(examples=examples, epochs=self.epochs_per_iteration, augment=True)
        self.iteration += 1
        if self.iteration % self.checkpoint_interval == 0 and _rank() == 0:      # data parallel: identical weights, one writer
            self.trainer.save_checkpoint(iteration=self.iteration)
        logger.info(f"Iteration {self.iteration} completed. Policy Loss: {metrics['policy_loss'][-1]:.4f}, "
                    f"Value Loss: {metrics['value_loss'][-1]:.4f}, Total Loss: {metrics['total_loss'][-1]:.4f}")
        return metrics

    def train(self, num_iterations=10):
        all_metrics = {"policy_loss": [], "value_loss": [], "total_loss": []}
        for _ in range(num_iterations):
            metrics = self.train_iteration()
            for key in all_metrics:
                all_metrics[key].extend(metrics.get(key, []))
        return all_metrics

    def get_latest_model_path(self):
        name = f"checkpoint_{self.iteration}.pth.tar" if self.iteration > 0 else "checkpoint.pth.tar"
        return os.path.join(self.model_dir, name)


def run_training_pipeline(game, model_dir="models", data_dir="data", num_iterations=10, sample_size=10000, checkpoint_interval=10,
                          **pipeline_kwargs):  # training_pipeline.py:289-330
    pipeline = TrainingPipeline(game=game, model_dir=model_dir, data_dir=data_dir, sample_size=sample_size,
                                checkpoint_interval=checkpoint_interval, **pipeline_kwargs)
    pipeline.load_data()
    pipeline.train(num_iterations=num_iterations)
    return pipeline.get_latest_model_path()
