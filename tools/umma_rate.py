"""Developer tool: tcgen05.mma issue rate vs shared-memory descriptor layout (cycles per M=128,N,K=16 MMA)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import yy_b200  # noqa
from yinyang_game_alphazero_b200 import _lib
L = _lib.lib()
L.yy_umma_rate.argtypes = [ctypes.c_int] * 11 + [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(148, dtype=torch.int64, device="cuda")
per_round, iters = 256, 20

def run(name, N, lt, lbo_a, sbo_a, lbo_b, sbo_b, a_step, b_step, ctas=148):
    rc = L.yy_umma_rate(N, lt, lbo_a, sbo_a, lbo_b, sbo_b, a_step, b_step, per_round, iters, ctas, ctypes.c_void_p(out.data_ptr()), None)
    assert rc == 0, L.yy_last_error()
    torch.cuda.synchronize()
    c = out[:ctas].float() / (per_round * iters)
    print(f"{name:64s} N={N:3d} cycles/MMA mean={c.mean().item():7.1f} min={c.min().item():7.1f} max={c.max().item():7.1f}", flush=True)

# no-swizzle K-major: LBO between k-chunks, SBO=128 between 8-row groups
run("none: A LBO 8960 (mod128=0), B LBO 2048 (mod128=0) [current]", 128, 0, 8960, 128, 2048, 128, 16, 4096)
run("none: A LBO 9024 (mod128=64), B LBO 2112 (mod128=64)", 128, 0, 9024, 128, 2112, 128, 16, 4224)
run("none: A LBO 9024, B LBO 2048", 128, 0, 9024, 128, 2048, 128, 16, 4096)
run("none: A LBO 8960, B LBO 2112", 128, 0, 8960, 128, 2112, 128, 16, 4224)
run("none: A LBO 8992 (mod128=32), B LBO 2080 (mod128=32)", 128, 0, 8992, 128, 2080, 128, 16, 4160)
run("none: A,B LBO=128 SBO=256 (k-chunks adjacent)", 128, 0, 128, 256, 128, 256, 32, 8192)
run("none: A aligned steps (a_step 128)", 128, 0, 8960, 128, 2048, 128, 128, 4096)
run("none N=256: A LBO 8960, B LBO 4096", 256, 0, 8960, 128, 4096, 128, 16, 8192)
run("none N=256: A LBO 9024, B LBO 4160", 256, 0, 9024, 128, 4160, 128, 16, 8320)
run("none N=64:  A LBO 8960, B LBO 1024", 64, 0, 8960, 128, 1024, 128, 16, 2048)
run("none N=64:  A LBO 9024, B LBO 1088", 64, 0, 9024, 128, 1088, 128, 16, 2176)
# 128B swizzle K-major: rows of 128 B, 8-row groups 1024 B apart; LBO ignored (1)
run("sw128: A,B SBO 1024, step 32 B (k16 within the 128B row)", 128, 2, 16, 1024, 16, 1024, 32, 32)
run("sw128 N=256", 256, 2, 16, 1024, 16, 1024, 32, 32)
run("sw128 N=64", 64, 2, 16, 1024, 16, 1024, 32, 32)
run("sw64: SBO 512, step 32", 128, 4, 16, 512, 16, 512, 32, 32)
run("sw32: SBO 256, step 32", 128, 6, 16, 256, 16, 256, 32, 32)
run("single CTA none current", 128, 0, 8960, 128, 2048, 128, 16, 4096, ctas=1)
