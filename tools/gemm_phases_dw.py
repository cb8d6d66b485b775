"""Developer tool: clock64 phase stamps of CTA 0 of the learner's WEIGHT-GRADIENT GEMM (dW[co][tap, ci] = sum_p dY[p][co] * X[p + d(tap)][ci]:
A = dY^T [128, 4096], B = the transposed activation [128, 4096] read as its transposed im2col, N = 1152, split-K 16)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import learner, _lib
ops = learner.CudaOps("3xtf32")
L = _lib.lib()
P = int(os.environ.get("YY_POSITIONS", 4096))
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
dYT, XT, dW = torch.randn(128, P).cuda(), torch.randn(128, P).cuda(), torch.zeros(128, 1152).cuda()


def run():
    ops.gemm(dYT, XT, dW, conv_t=(8, 8, 128, 0))


for _ in range(3):
    run()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(20):
    run()
ev1.record(); torch.cuda.synchronize()
print("us per weight-gradient GEMM (+ reducer)", ev0.elapsed_time(ev1) / 20 * 1e3)
p = lambda t: ctypes.c_void_p(t.data_ptr())
L.yy_lrn_gemm_debug_stamps(p(dbg))
run()
torch.cuda.synchronize()
L.yy_lrn_gemm_debug_stamps(None)
d = dbg.cpu().tolist()
t0 = d[0]
print("kernel entry", d[100] - t0, "predecessors done", d[103] - t0, "| setup", d[1] - t0, "loop end", d[2] - t0, "acc complete", d[119] - t0, "epilogue end", d[3] - t0,
      "cluster reduction end", d[101] - t0 if d[101] else None, "exit", d[102] - t0)
names = ["iter start", "next loads issued", "slot free", "stored + published"]
for k in range(8):
    print(k, dict(zip(names, [d[4 + 6 * k + j] - t0 if d[4 + 6 * k + j] else None for j in range(4)])))
