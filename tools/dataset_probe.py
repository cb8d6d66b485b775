"""Developer tool: the augmentation kernel alone (8x8, 262,144 records) -- timing by CUDA events and the target of
`ncu --set full -k regex:augment_kernel -c 1 python tools/dataset_probe.py`."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import engine

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
A = n * n
plies = torch.arange(n_rec, dtype=torch.int32) % (A - 12)
black, white, _ = engine.random_playout(n_rec, plies, n, n, seed=0xDA7A)
counts = torch.randint(0, 800, (n_rec, A), dtype=torch.int16, device="cuda")
values = torch.ones(n_rec, dtype=torch.float32, device="cuda")
out = engine.augment_samples(black, white, n, n, counts=counts, values=values)   # outputs allocated once, reused below
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
ev0.record()
for _ in range(reps):
    o = engine.augment_samples(black, white, n, n, counts=counts, values=values)
    del o
ev1.record(); torch.cuda.synchronize()
sec = ev0.elapsed_time(ev1) * 1e-3 / reps
W = (A + 63) // 64
bytes_per_record = 16 * W + 2 * A + 4 + 8 * (6 * A + 1) * 4
print(json.dumps({"board": n, "records": n_rec, "ms_per_launch": sec * 1e3, "GB/s": bytes_per_record * n_rec / sec / 1e9,
                  "samples/s": 8 * n_rec / sec}))
# the same launches through the C ABI on preallocated outputs (no allocator in the loop), one event pair per launch
from yinyang_game_alphazero_b200 import _lib
planes, pol, vals = out
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
args = (n, n, black.data_ptr(), white.data_ptr(), counts.data_ptr(), None, values.data_ptr(), n_rec, planes.data_ptr(), pol.data_ptr(), vals.data_ptr(), st)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
evs[0].record()
for k in range(10):
    _lib.check(L.yy_augment_samples(*args))
    evs[k + 1].record()
torch.cuda.synchronize()
per = [evs[k].elapsed_time(evs[k + 1]) for k in range(10)]
print(json.dumps({"preallocated_ms_per_launch": per, "GB/s_best": bytes_per_record * n_rec / (min(per) * 1e-3) / 1e9,
                  "GB/s_median": bytes_per_record * n_rec / (sorted(per)[5] * 1e-3) / 1e9}))
