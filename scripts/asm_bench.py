"""Times the C3 band assembly (V + T, k = 10, 2e5 exponential intervals) and C1 (k = 7) through the C ABI and checks
both against the oracle.  CB_ASM_STREAM_NS=<n> sets the step size of the streaming kernel (0 = tiled kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import compactbases_jl_b200 as cb
from compactbases_jl_b200 import _lib
from compactbases_jl_b200.operators import _ptr, _stream
from oracle import cbo

def relerr(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

for (K, ks, nint, check) in ((10, cb.ExpKnotSet(10, -3.0, 2.0, 200_000), 200_000, True), (7, cb.LinearKnotSet(7, 0, 100, 1000), 1000, True),
                             (8, cb.LinearKnotSet(8, 0, 500, 50_000), 50_000, False)):
    B = cb.BSpline(ks)
    W = 2 * K - 1
    V = torch.empty((B.nf, W), dtype=torch.float64, device="cuda"); T = torch.empty_like(V)
    st = _stream()
    def run():
        _lib.call("cb_bspline_assemble_coulomb", B.handle, B.handle, 1.0, 0, B.nf, 0, B.nf, K - 1, K - 1, _ptr(V), st)
        _lib.call("cb_bspline_assemble", B.handle, B.handle, 1, 1, None, 0, B.nf, 0, B.nf, K - 1, K - 1, _ptr(T), st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    V2 = torch.empty_like(V); T2 = torch.empty_like(T)
    def pair():
        _lib.call("cb_bspline_assemble_pair", B.handle, None, 1, 1.0, 0, B.nf, _ptr(V2), _ptr(T2), st)
    for _ in range(3): pair()
    torch.cuda.synchronize()
    tp = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pair(); e1.record(); torch.cuda.synchronize(); tp.append(e0.elapsed_time(e1) * 1e3)
    same = bool(torch.equal(V, V2) and torch.equal(T, T2))
    if not same:
        dv, dt = (V - V2).abs(), (T - T2).abs()
        print(f"  V: {int((dv > 0).sum())} of {dv.numel()} differ, max {float(dv.max()):.3e} (scale {float(V.abs().max()):.3e}); "
              f"T: {int((dt > 0).sum())} differ, max {float(dt.max()):.3e} (scale {float(T.abs().max()):.3e})", flush=True)
        iv = torch.nonzero(dv > 0)[:5].tolist(); it = torch.nonzero(dt > 0)[:5].tolist()
        print("  first V diffs", iv, "first T diffs", it, flush=True)
    msg = f"K={K} nint={nint} NS={os.environ.get('CB_ASM_STREAM_NS', 'default')}: V+T {np.mean(ts):.1f} us (min {np.min(ts):.1f}); pair {np.mean(tp):.1f} us (min {np.min(tp):.1f}) identical={same}"
    if check:
        cbo.set_threads(cbo.max_threads())
        te = cbo.expand_knots(np.asarray(ks.t), K); xq, wq = cbo.lgwt(te, K + 1)
        Vo, _, _ = cbo.bspline_assemble(te, K, te, K, xq, wq, K + 1, 0, 0, -1.0 / xq)
        To, _, _ = cbo.bspline_assemble(te, K, te, K, xq, wq, K + 1, 1, 1)
        msg += f"  rel err V {relerr(V.cpu().numpy(), Vo):.1e} T {relerr(T.cpu().numpy(), To):.1e}"
    print(msg, flush=True)
