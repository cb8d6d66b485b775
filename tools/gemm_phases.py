"""Developer tool: clock64 phase stamps of CTA 0 of one learner GEMM (yy_lrn_gemm_debug_stamps).
usage: gemm_phases.py [3xtf32|tf32] [split_k] [packed]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import learner, _lib
prec = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
split = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ops = learner.CudaOps(prec)
L = _lib.lib()
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
X, W, Y = torch.randn(4096, 128).cuda(), torch.randn(128, 1152).cuda(), torch.zeros(4096, 128).cuda()
geom = _lib.ConvGeom(8, 8, 128, 0)
packed = torch.empty(max(1, ops.packed_b_bytes(128, 1152)), dtype=torch.uint8, device="cuda")
use_packed = len(sys.argv) > 3 and sys.argv[3] == "packed" and ops.precision == 1
if use_packed:
    ops.pack_b(W, None, 0, 1, 128, 1152, packed)
bp = ctypes.c_void_p(packed.data_ptr()) if use_packed else None
p = lambda t: ctypes.c_void_p(t.data_ptr())


def run():
    _lib.check(L.yy_lrn_gemm(p(X), 128, 1, p(W), 1152, 0, p(Y), 128, 4096, 128, 1152, None, 0, 0, 128, split, p(ops.ws), ops.ws.numel(),
                             ops.precision, ctypes.byref(geom), None, bp, None))


for _ in range(3):
    run()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(20):
    run()
ev1.record(); torch.cuda.synchronize()
print(prec, "split", split, "us per GEMM (+ reducer)", ev0.elapsed_time(ev1) / 20 * 1e3)
L.yy_lrn_gemm_debug_stamps(p(dbg))
run()
torch.cuda.synchronize()
L.yy_lrn_gemm_debug_stamps(None)
d = dbg.cpu().tolist()
t0 = d[0]
print("kernel entry", d[100] - t0, "predecessors done", d[103] - t0, "| setup", d[1] - t0, "loop end", d[2] - t0, "acc complete", d[119] - t0, "epilogue end", d[3] - t0,
      "cluster reduction end", d[101] - t0 if d[101] else None, "exit", d[102] - t0)
names = ["iter start", "next loads issued", "slot free", "stored + published"]
for k in range(min(10, (1152 // split + 31) // 32)):
    print(k, dict(zip(names, [d[4 + 6 * k + j] - t0 if d[4 + 6 * k + j] else None for j in range(4)])))
