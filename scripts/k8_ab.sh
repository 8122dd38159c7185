#!/bin/bash
# A/B of apply-pipeline variants on the headline step, timing only (development; a variant may compute wrong results)
one() { python - "$1" <<'PY'
import sys, torch, numpy as np
import compactbases_jl_b200 as cb
F = cb.FEDVR(np.linspace(0, 1000, 10001), 10); D = cb.Derivative()
A = F.T @ D.T @ D @ F
X = torch.randn((1024, F.N), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
for _ in range(5): cb.mul(Y, A, X)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for rep in range(5):
    e0.record()
    for _ in range(20): cb.mul(Y, A, X)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 20 * 1e3)
print(sys.argv[1], "apply us:", " ".join(f"{t:.1f}" for t in ts))
PY
}
one default
for v in "$@"; do CB_CUDA_LIB=$PWD/variants/lib$v.so one $v; done
