"""Developer tool: only the env leg of bench.py (BASELINE.json configs[1])."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
import bench
from yinyang_game_alphazero_b200 import engine

print(json.dumps(bench.bench_env(engine, torch, bench.measured_peaks())))
