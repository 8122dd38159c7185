"""Times the pivoted band LU solve at the C5 basis size (50 007 x 1024, l = u = 7)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import compactbases_jl_b200 as cb
B5 = cb.BSpline(cb.LinearKnotSet(8, 0, 500, 50_000))
S5 = B5.T @ B5
F = cb.BandLU(S5)
X = torch.randn((1024, B5.nf), dtype=torch.float64, device="cuda")
for _ in range(2): F.ldiv(X); X.normal_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); F.ldiv(X); e1.record(); torch.cuda.synchronize()
print("bandlu solve ms", e0.elapsed_time(e1))
