"""Host-side mirror of the outer loop, ``AlphaZero`` (src/yin_yang/ai/alphazero.py:21-270): per iteration self-play with the
best model -> train the current model on the data directory -> evaluate current vs best -> promote at a win ratio >= 0.6.
Every stage runs on the GPU: batched self-play (self_play.generate_self_play_data), the CUDA learner
(training_pipeline.run_training_pipeline) and the batched arena (arena.evaluate)."""
from __future__ import annotations

import logging
import os
import shutil

from . import arena
from .network import YinYangNeuralNetwork
from .self_play import generate_self_play_data
from .training_pipeline import run_training_pipeline

logger = logging.getLogger("YinYangAlphaZero")


class AlphaZero:
    def __init__(self, game, model_dir="models", data_dir="data", num_iterations=100, num_episodes=100, num_simulations=800,
                 num_epochs=10, temperature_threshold=10, update_threshold=0.6, num_workers=1, mcts_threads=1, eval_games=40,
                 batch_size=64, lr=0.001):
        self.game, self.model_dir, self.data_dir = game, model_dir, data_dir
        self.num_iterations, self.num_episodes, self.num_simulations = num_iterations, num_episodes, num_simulations
        self.num_epochs, self.temperature_threshold, self.update_threshold = num_epochs, temperature_threshold, update_threshold
        self.num_workers, self.mcts_threads, self.eval_games = num_workers, mcts_threads, eval_games
        self.batch_size, self.lr = batch_size, lr
        for d in (model_dir, data_dir):
            os.makedirs(d, exist_ok=True)
        self.current_model_path = os.path.join(model_dir, "current_model.pth.tar")
        self.best_model_path = os.path.join(model_dir, "best_model.pth.tar")
        if not os.path.exists(self.current_model_path):
            YinYangNeuralNetwork(game).save_model(self.current_model_path)          # alphazero.py:79-83
        if not os.path.exists(self.best_model_path):
            shutil.copy(self.current_model_path, self.best_model_path)

    def self_play(self, model_path):  # alphazero.py:85-108
        return generate_self_play_data(game=self.game, model_path=model_path, output_dir=self.data_dir, num_games=self.num_episodes,
                                       num_workers=self.num_workers, num_simulations=self.num_simulations)

    def train(self):  # alphazero.py:110-134: one pipeline iteration over everything in data_dir, checkpoint -> current model
        new_model_path = run_training_pipeline(game=self.game, model_dir=self.model_dir, data_dir=self.data_dir, num_iterations=1,
                                               sample_size=10000, checkpoint_interval=1, epochs_per_iteration=self.num_epochs,
                                               batch_size=self.batch_size, lr=self.lr)
        shutil.copy(new_model_path, self.current_model_path)
        return self.current_model_path

    def evaluate(self, current_model_path, best_model_path, num_games=None):  # alphazero.py:136-226
        return arena.evaluate(self.game, current_model_path, best_model_path, num_games=num_games or self.eval_games,
                              num_simulations=self.num_simulations, mcts_threads=self.mcts_threads)

    def update_best_model(self, win_ratio):  # alphazero.py:228-246
        if arena.should_promote(win_ratio, self.update_threshold):
            shutil.copy(self.current_model_path, self.best_model_path)
            logger.info(f"Updated best model with win ratio {win_ratio:.2f} >= {self.update_threshold}")
            return True
        logger.info(f"Kept best model with win ratio {win_ratio:.2f} < {self.update_threshold}")
        return False

    def run(self):  # alphazero.py:248-270
        for iteration in range(self.num_iterations):
            logger.info(f"Starting iteration {iteration + 1}/{self.num_iterations}")
            self.self_play(self.best_model_path)
            self.train()
            win_ratio = self.evaluate(self.current_model_path, self.best_model_path)
            self.update_best_model(win_ratio)
            logger.info(f"Completed iteration {iteration + 1}/{self.num_iterations}")
