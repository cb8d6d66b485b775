"""Developer tool: only the learner leg of bench.py."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
import bench
from yinyang_game_alphazero_b200 import engine

print(json.dumps(bench.bench_learner(engine, torch, bench.measured_peaks())))
