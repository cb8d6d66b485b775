"""Developer tool: clock64 phase stamps of CTA 0 of one learner GEMM (yy_lrn_gemm_debug_stamps)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import learner, _lib
prec = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
ops = learner.CudaOps(prec)
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
X, W, Y = torch.randn(4096, 128).cuda(), torch.randn(128, 1152).cuda(), torch.zeros(4096, 128).cuda()
for _ in range(3):
    ops.gemm(X, W, Y, conv=(8, 8, 128, 0))
_lib.lib().yy_lrn_gemm_debug_stamps(ctypes.c_void_p(dbg.data_ptr()))
ops.gemm(X, W, Y, conv=(8, 8, 128, 0))
torch.cuda.synchronize()
_lib.lib().yy_lrn_gemm_debug_stamps(None)
d = dbg.cpu().tolist()
t0 = d[0]
print(prec, "setup", d[1] - t0, "loop end", d[2] - t0, "acc complete", d[119] - t0, "epilogue end", d[3] - t0)
names = ["iter start", "next loads issued", "slot free", "stored + published", "-", "-"]
for k in range(9):
    row = [d[4 + 6 * k + j] - t0 if d[4 + 6 * k + j] else None for j in range(6)]
    print(k, dict(zip(names, row)))
