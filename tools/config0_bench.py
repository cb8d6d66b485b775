"""Developer tool: the configs[0] leg of bench.py alone (6x6, 100 simulations, 4,096 rolling games)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
import bench
from yinyang_game_alphazero_b200 import engine, network
print(json.dumps(bench.bench_config(engine, network, torch, bench.measured_peaks(), 6, 100, 4096, steps=3, iters=8 * 101)))
