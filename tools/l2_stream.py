"""Developer tool: L2 -> shared-memory streaming rate with cp.async.bulk when all SMs pull the same buffer."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import _lib
L = _lib.lib()
L.yy_l2_stream.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
src = torch.randint(0, 255, (7 * 1024 * 1024 + 512 * 1024,), dtype=torch.uint8, device="cuda")
out = torch.zeros(148, dtype=torch.int64, device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for chunk, slots, warps, ctas in [(16384, 5, 1, 148), (16384, 2, 2, 148), (16384, 2, 4, 148), (8192, 3, 4, 148), (8192, 2, 8, 148), (4096, 4, 8, 148), (16384, 1, 4, 148), (32768, 2, 2, 148)]:
    total = 64 * 7 * 1024 * 1024 // 8 if ctas > 1 else 7 * 1024 * 1024 * 4
    total = (total // (chunk * warps)) * chunk * warps
    for rep in range(2):
        ev0.record()
        rc = L.yy_l2_stream(ctypes.c_void_p(src.data_ptr()), 7 * 1024 * 1024, total, chunk, slots, warps, ctas, ctypes.c_void_p(out.data_ptr()), None)
        ev1.record(); torch.cuda.synchronize()
        assert rc == 0, L.yy_last_error()
    ms = ev0.elapsed_time(ev1)
    cyc = out[:ctas].float()
    print(f"chunk={chunk} slots={slots} warps={warps} ctas={ctas}: {total/cyc.mean().item():.1f} B/clk/SM  total {total*ctas/ms/1e6:.0f} GB/s  ({ms:.2f} ms)", flush=True)
