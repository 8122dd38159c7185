"""Times B[x,:] as SparseMatrixCSC (cb_bspline_eval_csc, preallocated outputs) at the C1 size."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import compactbases_jl_b200 as cb
from compactbases_jl_b200 import _lib
from compactbases_jl_b200.operators import _ptr, _stream
B = cb.BSpline(cb.LinearKnotSet(7, 0, 100, 1000))
for NX, kind in ((1_000_000, "random"), (1_000_000, "sorted"), (16_000_000, "random")):
    x = torch.rand(NX, dtype=torch.float64, device="cuda") * 100
    if kind == "sorted": x = torch.sort(x).values
    colptr = torch.empty(B.nf + 1, dtype=torch.int64, device="cuda")
    rowval = torch.empty(NX * 7, dtype=torch.int64, device="cuda"); nzval = torch.empty(NX * 7, dtype=torch.float64, device="cuda")
    nnz = C.c_int64(0)
    fn = lambda: _lib.call("cb_bspline_eval_csc", B.handle, _ptr(x), NX, 0, B.nf, _ptr(colptr), _ptr(rowval), _ptr(nzval), C.byref(nnz), _stream())
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(f"csc {kind} nx={NX}: {np.mean(ts)*1e6:.1f} us (min {np.min(ts)*1e6:.1f}) nnz={nnz.value}", flush=True)
    del x, rowval, nzval
