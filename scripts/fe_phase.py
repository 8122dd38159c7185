"""Phase-time breakdown of the lean FE-DVR apply kernel (variants/libfe_prof.so, built with -DFEDVR_PHASE_PROF)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CB_CUDA_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "variants", "libfe_prof.so"))
import numpy as np, torch
import compactbases_jl_b200 as cb
lib = C.CDLL(os.environ["CB_CUDA_LIB"])
D = cb.Derivative()
F = cb.FEDVR(np.linspace(0, 1000, 10001), 10)
A = F.T @ D.T @ D @ F
X = torch.randn((1024, F.N), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
for _ in range(3): cb.mul(Y, A, X)
torch.cuda.synchronize()
out = (C.c_ulonglong * 8)()
lib.cb_debug_fedvr_prof(out, 1)
N = 5
for _ in range(N): cb.mul(Y, A, X)
torch.cuda.synchronize()
lib.cb_debug_fedvr_prof(out, 0)
names_pipe = ["L: wait empty", "L: issue copies", "C: wait full", "C: products+writeback", "C: named barrier", "S: wait done", "S: store", "C: bridge+frag loads"]
names = ["issue loads+frag", "cp.async wait", "barrier(load)", "compute", "barrier(compute)", "bridge+barrier", "store issue", "barrier(store)"]
tiles = 667 * 32 * N
tot = sum(out)
for n, v in zip(names_pipe if os.environ.get("FE_PIPE") else names, out):
    print(f"{n:20s} {v / tiles:9.0f} cycles/tile  {100 * v / tot:5.1f}%")
print("total cycles/tile", tot / tiles)
