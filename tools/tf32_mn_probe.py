"""Developer tool: does tcgen05.mma kind::tf32 accept a transposed (MN-major) A operand, and in which shared-memory layout?"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import _lib
L = _lib.lib()
p = lambda t: ctypes.c_void_p(t.data_ptr())
torch.manual_seed(0)
for K in (8, 32):
    for N in (128, 32):
        A = torch.randn(128, K); B = torch.randn(N, K)
        ref = A.double() @ B.double().t()
        At = A.t().contiguous().cuda(); Bd = B.cuda()
        for variant in (0, 1):
            C = torch.full((128, N), -7.0).cuda()
            _lib.check(L.yy_probe_tf32_mn(p(At), p(Bd), p(C), N, K, variant, None))
            torch.cuda.synchronize()
            err = (C.cpu().double() - ref).abs().max().item()
            print(f"K={K} N={N} variant={variant}: max err {err:.3e}  nonzero {int((C != 0).sum())}  C[0,:3]={C[0,:3].tolist()} ref={ref[0,:3].tolist()}")
