"""Phase-time breakdown of the pipelined banded apply kernel (variants/libbt_prof.so, built with -DFEDVR_PHASE_PROF)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CB_CUDA_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "variants", "libbt_prof.so"))
import numpy as np, torch
import compactbases_jl_b200 as cb
lib = C.CDLL(os.environ["CB_CUDA_LIB"])
D = cb.Derivative()
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    B = cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, 200000)); nv = 256
else:
    B = cb.BSpline(cb.LinearKnotSet(7, 0, 100, 1000)); nv = 65536
S = B.T @ D.T @ D @ B
X = torch.randn((nv, B.nf), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
for _ in range(3): cb.mul(Y, S, X)
torch.cuda.synchronize()
out = (C.c_ulonglong * 8)()
lib.cb_debug_fedvr_prof(out, 1)
N = 5
for _ in range(N): cb.mul(Y, S, X)
torch.cuda.synchronize()
lib.cb_debug_fedvr_prof(out, 0)
names = ["L: wait empty", "L: issue copies", "C: wait full", "C: products", "C: named barrier", "S: wait done", "S: store", "C: writeback+arrive (+A staging)"]
tiles = ((B.nf + 127) // 128) * ((nv + 31) // 32) * N
tot = sum(out)
for n, v in zip(names, out):
    print(f"{n:34s} {v / tiles:9.0f} cycles/tile  {100 * v / tot:5.1f}%")
print("total cycles/tile/role", tot / tiles / 3)
