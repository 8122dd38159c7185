#!/usr/bin/env python
"""bench.py — benchmark of the hot path (contract: see DESIGN.md §6).

Headline workload at N GPUs (BASELINE.json configs[1], "C2"): FE-DVR order 10 with 1e4 elements
(n = 90 001 functions); one STEP = assemble the block-banded second-derivative operator B'D'DB (K4) and
apply it to a batch of 1024 coefficient vectors (K8) through ONE foreign call (cb_fedvr_assemble_apply),
all Float64.  The vector batch is the sharded dimension: every rank holds its own 1024 vectors and its
own copy of the operator (no data-path collective, weak scaling).

  value     algorithmic GB/s of the whole job with inputs resident in HBM
            (bytes/step = read X + write Y + element blocks written once and read once)
  e2e       the same metric through the host-buffer C-ABI entry (cb_fedvr_apply_host): pinned host
            X -> device -> Y back to the host inside the timed region
  roofline  the apply kernel alone (CUDA events around its launches) vs the measured HBM peak
  cpu_baseline / --impl reference   the CPU restatement of the reference (oracle/), single thread and
            OpenMP over all host cores, on a bounded slice of the same workload

"configs" carries the other BASELINE configurations — C1 (B-spline evaluation + small assembly), C3 (band
assembly on 2e5 exponential intervals), C4 (tridiagonal apply, N = 1e7 x 4096 vectors), C5 (density
projection, 1e4 pairs) — each with its own value, roofline, e2e (host buffers through the C ABI),
cpu_baseline (N = 1 only) and a parity check of the timed output against the oracle.  These are STRONG
scaling runs: the job is fixed and its points / intervals / vectors / pairs are split over the ranks
(SURVEY §8e), timed on the device as the max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NEL, ORDER, NVEC = 10_000, 10, 1024
N_FUN = NEL * (ORDER - 1) + 1
BYTES_APPLY = 2 * 8 * N_FUN * NVEC + 8 * NEL * ORDER * ORDER          # X + Y + blocks read
BYTES_ASSEMBLE = 8 * NEL * ORDER * ORDER + 8 * (2 * N_FUN + NEL * ORDER)  # blocks written + x, n, w read
BYTES_STEP = BYTES_APPLY + BYTES_ASSEMBLE
METRIC = "fedvr_assemble_apply_algorithmic_GBps"
WORKLOAD = "C2: FEDVR order 10, 1e4 elements; assemble B'D'DB + apply to 1024 Float64 vectors per GPU"
RTOL = 1e-12            # north_star: <= 1e-12 relative in Float64


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = float(np.max(np.abs(b))) if b.size else 0.0
    return float(np.max(np.abs(a - b))) / (den if den > 0 else 1.0) if a.size else 0.0


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# --------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference algorithms (bench.py's cpu_baseline leg and
# --impl reference are the only places outside tests/ and smoke() that execute oracle/)
# --------------------------------------------------------------------------------------
_CPU_STATE = {}


_ALL_CPUS = None         # affinity of the process before bench.py bound it to the GPU's NUMA node


def _cbo(threads=True):
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm may use every host core
    if _ALL_CPUS is not None:
        try:
            os.sched_setaffinity(0, _ALL_CPUS)              # the CPU legs are not confined to the GPU's NUMA node
        except OSError:
            pass
    if "oracle.cbo" not in sys.modules:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import cbo
    cbo.set_threads(cbo.max_threads() if threads else 1)
    return cbo


def _best(fn, repeats):
    best = float("inf")
    out = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def cpu_reference_pass(nvec_sample, repeats=3, threads=True):
    cbo = _cbo(threads)
    if nvec_sample not in _CPU_STATE:                       # basis construction + inputs are outside the timed region
        t = np.linspace(0.0, 1000.0, NEL + 1)
        rng = np.random.default_rng(2)
        _CPU_STATE[nvec_sample] = (cbo.FEDVROracle(t, ORDER), np.asfortranarray(rng.standard_normal((N_FUN, nvec_sample))),
                                   np.zeros((N_FUN, nvec_sample), order="F"))
    O, X, Y = _CPU_STATE[nvec_sample]

    def one():
        blk = O.blocks(2)                                   # assemble (single-threaded, like the reference)
        O.apply(blk, X, 1.0, 0.0, Y, threads=threads)       # apply over the vector slice
    best, _ = _best(one, repeats)
    bytes_sample = 2 * 8 * N_FUN * nvec_sample + 8 * NEL * ORDER * ORDER + BYTES_ASSEMBLE
    cores = int(cbo.max_threads()) if threads else 1
    return bytes_sample / best / 1e9, best, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nvs = 256
    for _ in range(args.warmup):
        cpu_reference_pass(nvs, repeats=1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, sec, cores = cpu_reference_pass(nvs, repeats=1)
    total = time.perf_counter() - t0
    bytes_sample = 2 * 8 * N_FUN * nvs + 8 * NEL * ORDER * ORDER + BYTES_ASSEMBLE
    value = bytes_sample * args.steps / total / 1e9
    v1, _, _ = cpu_reference_pass(32, repeats=2, threads=False)   # the reference itself is single-threaded
    sample = f"assemble + apply on {nvs} of the {NVEC} vectors per step (CPU restatement of the reference, OpenMP)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample,
                         "single_thread_value": v1, "single_thread_sample": "32 of the 1024 vectors, best of 2"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------
# GPU arm: helpers
# --------------------------------------------------------------------------------------
class Ctx:
    """What every config needs: torch, the package, the rank layout, peaks and max-over-ranks reductions."""

    def __init__(self, torch, dist, cb, rank, world, peaks, peak_src, fp64_peak):
        self.torch, self.dist, self.cb, self.rank, self.world = torch, dist, cb, rank, world
        self.peak_gbs = float(peaks["hbm_gbs"])
        self.peak_src = peak_src + " (MEASURED_PEAKS.json hbm_gbs)"
        self.fp64_peak = fp64_peak           # TFLOP/s, DFMA microbenchmark on this GPU (libcb_micro.so)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxr(self, v):
        """max over ranks of a host scalar (device times of one rank each)."""
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def dev_time(self, fn, n, warm=2, rotate=1):
        """Device time (s) of fn(i % rotate): (mean, min).  mean = one CUDA event pair around n back-to-back launches on
        torch's current stream (the stream the library launches on) / n, as SURVEY 8d prescribes — the queue never runs
        empty, so no host latency is measured; min = the shortest of n launches timed one by one.  Barrier before each
        loop; both are the max over ranks."""
        torch = self.torch
        for i in range(warm):
            fn(i % rotate)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i % rotate)
        e1.record()
        torch.cuda.synchronize()
        mean = e0.elapsed_time(e1) * 1e-3 / n
        self.barrier()
        evs = []
        for i in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(i % rotate); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        mn = min(a.elapsed_time(b) for a, b in evs) * 1e-3
        return self.maxr(mean), self.maxr(mn)

    def wall_time(self, fn, n, warm=1):
        """Host wall clock (s per call) of a synchronous host-buffer entry, barriers on both sides, max over ranks."""
        for _ in range(warm):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        self.barrier()
        return self.maxr((time.perf_counter() - t0) / n)

    def hbm_roofline(self, bytes_per_launch, sec, kernel, note=None):
        ach = bytes_per_launch / sec / 1e9
        r = {"bound": "hbm", "achieved": ach, "peak": self.peak_gbs, "unit": "GB/s", "frac": ach / self.peak_gbs, "traffic": None,
             "kernel": kernel, "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms": sec * 1e3, "peak_source": self.peak_src}
        if note:
            r["note"] = note
        return r

    def fp64_roofline(self, flops_per_launch, bytes_per_launch, sec, kernel):
        ach = flops_per_launch / sec / 1e12
        return {"bound": "fp64", "achieved": ach, "peak": self.fp64_peak, "unit": "TFLOP/s", "frac": ach / self.fp64_peak,
                "traffic": None, "kernel": kernel, "algorithmic_flops_per_launch": flops_per_launch, "kernel_ms": sec * 1e3,
                "hbm_frac": bytes_per_launch / sec / 1e9 / self.peak_gbs, "algorithmic_bytes_per_launch": bytes_per_launch,
                "peak_source": "DFMA microbenchmark on this GPU in this run (libcb_micro.so: 8 independent chains, 2048 threads/SM)"}


def pinned(torch, shape, dtype=None):
    from compactbases_jl_b200.numa import bind_to_gpu_numa
    bind_to_gpu_numa(torch.cuda.current_device())          # (a CPU leg may have widened the affinity again)
    return torch.empty(shape, dtype=dtype or torch.float64, pin_memory=True)


def hptr(t):
    return C.c_void_p(t.data_ptr())


# --------------------------------------------------------------------------------------
# C1: B-spline k = 7 on LinearKnotSet(7, 0, 100, 1000): B[x,:] at 1e6 points (points split over the ranks),
#     mass + Laplacian assembly
# --------------------------------------------------------------------------------------
def config_c1(cx: Ctx, cpu: bool):
    torch, cb = cx.torch, cx.cb
    from compactbases_jl_b200 import _lib
    from compactbases_jl_b200.operators import _ptr, _stream
    from compactbases_jl_b200.sharding import batch_range
    K, NINT = 7, 1000
    t = cb.LinearKnotSet(K, 0, 100, NINT)
    B = cb.BSpline(t)
    D = cb.Derivative()
    out = {"workload": "C1: BSpline k=7, LinearKnotSet 0..100, 1000 intervals; B[x,:] at 1e6 unsorted points (68 B/point), "
                       "mass + Laplacian bands", "scaling": "strong", "sharding": "contiguous point ranges, knots replicated, no collective"}
    res = {}
    for NX, tag, reps in ((1_000_000, "eval_1e6", 40), (16_000_000, "eval_1.6e7", 10)):
        lo, hi = batch_range(NX, cx.rank, cx.world)
        nx = hi - lo
        # inputs of consecutive launches rotate over buffer sets whose total exceeds the 126 MB L2
        nset = max(1, -(-3 * 126_000_000 // (68 * nx))) if 68 * nx < 3 * 126_000_000 else 1
        rng = np.random.default_rng(1)
        xall = rng.uniform(0.0, 100.0, NX)[lo:hi]
        xs = [torch.from_numpy(np.roll(xall, 977 * s)).cuda() for s in range(nset)]
        ivs = [torch.empty(nx, dtype=torch.int32, device="cuda") for _ in range(nset)]
        vls = [torch.empty((nx, K), dtype=torch.float64, device="cuda") for _ in range(nset)]
        st = _stream()
        fn = lambda s: _lib.call("cb_bspline_eval", B.handle, _ptr(xs[s]), nx, 0, _ptr(ivs[s]), _ptr(vls[s]), st)
        mean, mn = cx.dev_time(fn, reps, warm=nset + 1, rotate=nset)
        r = {"points": NX, "us": mean * 1e6, "us_min": mn * 1e6, "evals_per_s": K * NX / mean, "points_per_s": NX / mean,
             "l2": f"{nset} rotating input/output sets ({68 * nx * nset / 1e6:.0f} MB > L2)" if nset > 1 else "inputs larger than L2",
             "roofline": cx.hbm_roofline(68 * nx, mean, "bspline_eval_kernel<7,0>",
                                         "per rank: 8 B x + 4 B interval + 56 B values per point")}
        if NX == 1_000_000 and cx.rank == 0:
            cbo = _cbo(True)
            te = cbo.expand_knots(t.t, K)
            ne = min(nx, 20000)
            edge = np.concatenate([xall[:ne], t.t, [np.nextafter(0.0, -1), np.nextafter(100.0, 200.0)]])
            iv, vals = B.eval_ell(edge)
            ivo, valo = cbo.bspline_eval_ell(te, K, K, edge)
            dv = np.abs(vals.cpu().numpy() - valo)
            r["parity"] = {"vs": "oracle bspline_eval_ell", "points": int(edge.size), "interval_exact": bool(np.array_equal(iv.cpu().numpy(), ivo)),
                           "max_elementwise_abs_over_scale": float(np.max(dv / np.maximum(np.abs(valo), 1e-300) * (np.abs(valo) > 1e-200))),
                           "max_rel": relerr(vals.cpu().numpy(), valo), "timed_output_slice_max_rel": relerr(vls[0][:ne].cpu().numpy(), valo[:ne])}
            assert r["parity"]["interval_exact"] and r["parity"]["max_rel"] <= RTOL and r["parity"]["timed_output_slice_max_rel"] <= RTOL, r["parity"]
        res[tag] = r
        del xs, ivs, vls
    out.update({"metric": "bspline_evals_per_s", "unit": "evals/s", "value": res["eval_1e6"]["evals_per_s"], "us": res["eval_1e6"]["us"],
                "roofline": res["eval_1e6"]["roofline"], "parity": res["eval_1e6"].get("parity"), "eval_1.6e7": res["eval_1.6e7"]})

    # ---- e2e: host x -> B[x,:] rows on the host (cb_bspline_eval_host; pinned buffers) ----
    NX = 1_000_000
    lo, hi = batch_range(NX, cx.rank, cx.world)
    nx = hi - lo
    xh, ivh, vh = pinned(torch, nx), pinned(torch, nx, torch.int32), pinned(torch, (nx, K))
    xh.copy_(torch.from_numpy(np.random.default_rng(1).uniform(0.0, 100.0, NX)[lo:hi]))
    sec = cx.wall_time(lambda: _lib.call("cb_bspline_eval_host", B.handle, hptr(xh), nx, 0, hptr(ivh), hptr(vh)), 10, warm=2)
    out["e2e"] = {"value": K * NX / sec, "unit": "evals/s", "us": sec * 1e6, "h2d_bytes_per_step": 8 * nx, "d2h_bytes_per_step": 60 * nx,
                  "entry": "cb_bspline_eval_host (pinned host x, interval, vals)", "GBps": 68 * NX / sec / 1e9}

    # ---- B[x,:] as Julia's SparseMatrixCSC through the API (colptr/rowval Int64, zeros dropped) ----
    if cx.rank == 0:
        x = torch.rand(NX, dtype=torch.float64, device="cuda") * 100
        colptr = torch.empty(B.nf + 1, dtype=torch.int64, device="cuda")
        rowval = torch.empty(NX * K, dtype=torch.int64, device="cuda")
        nzval = torch.empty(NX * K, dtype=torch.float64, device="cuda")
        nnz = C.c_int64(0)
        st = _stream()
        fn = lambda s: _lib.call("cb_bspline_eval_csc", B.handle, _ptr(x), NX, 0, B.nf, _ptr(colptr), _ptr(rowval), _ptr(nzval), C.byref(nnz), st)
        for _ in range(2):
            fn(0)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            t0 = time.perf_counter(); fn(0); ts.append(time.perf_counter() - t0)     # the call synchronises (nnz read-back)
        sec = float(np.mean(ts))
        byt = (8 + 16 * K) * NX + 8 * (B.nf + 1)
        out["eval_csc_1e6"] = {"us": sec * 1e6, "points_per_s": NX / sec, "GBps": byt / sec / 1e9, "hbm_frac": byt / sec / 1e9 / cx.peak_gbs,
                               "algorithmic_bytes": byt, "timing": "host wall clock of the synchronous call (preallocated outputs)", "nnz": int(nnz.value)}
        del x, colptr, rowval, nzval

    # ---- assembly: mass + Laplacian (launch-latency bound, not a roofline row); one basis and 64 bases back to back ----
    if cx.rank == 0:
        S = torch.empty((B.nf, 2 * K - 1), dtype=torch.float64, device="cuda")
        T = torch.empty_like(S)
        st = _stream()

        def asm(Bh, S, T):
            _lib.call("cb_bspline_assemble", Bh, Bh, 0, 0, None, 0, B.nf, 0, B.nf, K - 1, K - 1, _ptr(S), st)
            _lib.call("cb_bspline_assemble", Bh, Bh, 1, 1, None, 0, B.nf, 0, B.nf, K - 1, K - 1, _ptr(T), st)
        mean1, mn1 = cx_dev_time_local(torch, lambda: asm(B.handle, S, T), 50)
        meanp, mnp = cx_dev_time_local(torch, lambda: _lib.call("cb_bspline_assemble_pair", B.handle, None, 0, 0.0, 0, B.nf, _ptr(S), _ptr(T), st), 50)
        bases = [cb.BSpline(cb.LinearKnotSet(K, 0, 100 + i, NINT)) for i in range(64)]
        Ss = torch.empty((64, B.nf, 2 * K - 1), dtype=torch.float64, device="cuda")
        Ts = torch.empty_like(Ss)

        def asm64():
            for i, Bi in enumerate(bases):
                asm(Bi.handle, Ss[i], Ts[i])
        mean64, mn64 = cx_dev_time_local(torch, asm64, 10)
        api, _ = cx_dev_time_local(torch, lambda: (B.T @ B, B.T @ D.T @ D @ B), 20)
        cbo = _cbo(True)
        te = cbo.expand_knots(t.t, K)
        q = cbo.num_quadrature_points(K, 3)
        xq, wq = cbo.lgwt(te, q)
        So, _, _ = cbo.bspline_assemble(te, K, te, K, xq, wq, q, 0, 0)
        To, _, _ = cbo.bspline_assemble(te, K, te, K, xq, wq, q, 1, 1)
        par = {"vs": "oracle bspline_assemble", "mass_max_rel": relerr(S.cpu().numpy(), So), "laplacian_max_rel": relerr(T.cpu().numpy(), To)}
        assert max(par["mass_max_rel"], par["laplacian_max_rel"]) <= RTOL, par
        hostT = time.perf_counter()
        for _ in range(5):
            from compactbases_jl_b200.bsplines import diffop_host
            diffop_host(B, B, 0, 0); diffop_host(B, B, 1, 1)
        e2e_s = (time.perf_counter() - hostT) / 5
        out["assemble"] = {"us_per_mass_plus_laplacian": mean1 * 1e6, "us_min": mn1 * 1e6, "us_through_operator_api": api * 1e6,
                           "us_pair_entry": meanp * 1e6, "us_pair_entry_min": mnp * 1e6,
                           "batched_64_bases_us": mean64 * 1e6, "bases_per_s_batched": 64 / mean64, "us_per_basis_batched": mean64 / 64 * 1e6,
                           "note": "launch-latency bound (0.22 MB per basis); 2 launches per basis", "parity": par,
                           "e2e": {"us": e2e_s * 1e6, "entry": "cb_bspline_assemble_host x2 (bands to host arrays)",
                                   "d2h_bytes_per_step": 2 * 8 * B.nf * (2 * K - 1), "h2d_bytes_per_step": 0}}
        del bases, Ss, Ts

    # ---- CPU: single thread, OpenMP, and the reference's literal O(nf k nx) scan at 1e4 points ----
    if cpu and cx.rank == 0:
        cbo = _cbo(False)
        te = cbo.expand_knots(t.t, K)
        x = np.random.default_rng(1).uniform(0.0, 100.0, NX)
        cbo.bspline_eval_ell(te, K, K, x[:1000])
        s1, _ = _best(lambda: cbo.bspline_eval_ell(te, K, K, x), 2)
        cbo = _cbo(True)
        sN, _ = _best(lambda: cbo.bspline_eval_ell(te, K, K, x, threads=True), 3)
        sl, _ = _best(lambda: cbo.bspline_eval_literal(t.t, K, K, K, x[:10000]), 1)
        q = cbo.num_quadrature_points(K, 3)
        xq, wq = cbo.lgwt(te, q)
        cbo.set_threads(1)
        sa, _ = _best(lambda: (cbo.bspline_assemble(te, K, te, K, xq, wq, q, 0, 0), cbo.bspline_assemble(te, K, te, K, xq, wq, q, 1, 1)), 3)
        out["cpu_baseline"] = {"value": K * NX / sN, "unit": "evals/s", "cores": int(cbo.max_threads()), "kind": "port",
                               "sample": "all 1e6 points, linear-cost row form (bisection + de Boor), OpenMP over points, best of 3",
                               "single_thread_value": K * NX / s1, "single_thread_sample": "all 1e6 points, best of 2",
                               "literal_algorithm": {"value": K * 10000 / sl, "unit": "evals/s", "cores": 1,
                                                     "sample": "the reference's own getindex loop (every function x every point, "
                                                               "src/bsplines.jl:208-224) at 1e4 points"},
                               "assemble_mass_plus_laplacian_us_single_thread": sa * 1e6}
    return out


def cx_dev_time_local(torch, fn, n, warm=3):
    """Device time of fn() on this rank only (configs that rank 0 runs alone): (mean over n back-to-back launches
    between one event pair, min over n launches timed one by one)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    mean = e0.elapsed_time(e1) * 1e-3 / n
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(mean), float(min(a.elapsed_time(b) for a, b in evs) * 1e-3)


# --------------------------------------------------------------------------------------
# C3: BSpline k = 10 on ExpKnotSet(10, -3, 2, 2e5): Coulomb potential + Laplacian bands, intervals sharded
# --------------------------------------------------------------------------------------
C3_FLOP = 1.6e9          # 0.9 GFLOP contraction + 0.7 GFLOP basis tables for V + T at 2e5 intervals
C3_BYTES = 62.4e6        # knots in + two 19 x 200 009 bands out


def config_c3(cx: Ctx, cpu: bool, comm):
    torch, cb = cx.torch, cx.cb
    from compactbases_jl_b200 import _lib
    from compactbases_jl_b200.bsplines import diffop_host
    from compactbases_jl_b200.operators import _ptr, _stream
    from compactbases_jl_b200.sharding import assemble_interval_sharded
    K = 10
    D = cb.Derivative()
    out = {"workload": "C3: BSpline k=10, ExpKnotSet(10,-3,2,200000) (q=11): Coulomb potential V = -1/r and Laplacian T bands "
                       "(2 x 19 x 200 009)", "scaling": "strong",
           "sharding": "contiguous t.t interval ranges; every rank fills the band columns it owns; the k-1 boundary columns are "
                       "completed by the fused NVLink mailbox exchange (halo=fused), an NCCL send/recv (halo=exchange) or by "
                       "recomputing the boundary intervals (halo=recompute)"}
    sizes = {}
    for nint, tag in ((200_000, "2e5"), (1_600_000, "1.6e6")):
        B = cb.BSpline(cb.ExpKnotSet(K, -3.0, 2.0, nint))
        scale = nint / 200_000
        W = 2 * K - 1
        st = _stream()
        full_V = torch.empty((B.nf, W), dtype=torch.float64, device="cuda")
        full_T = torch.empty_like(full_V)

        def single():
            _lib.call("cb_bspline_assemble_coulomb", B.handle, B.handle, 1.0, 0, B.nf, 0, B.nf, K - 1, K - 1, _ptr(full_V), st)
            _lib.call("cb_bspline_assemble", B.handle, B.handle, 1, 1, None, 0, B.nf, 0, B.nf, K - 1, K - 1, _ptr(full_T), st)
        s_mean, s_min = cx_dev_time_local(torch, single, 10)
        p_mean, p_min = cx_dev_time_local(torch, lambda: _lib.call("cb_bspline_assemble_pair", B.handle, None, 1, 1.0, 0, B.nf,
                                                                   _ptr(full_V), _ptr(full_T), st), 10)
        r = {"intervals": nint, "single_gpu_us": s_mean * 1e6, "single_gpu_us_min": s_min * 1e6, "pair_entry_us": p_mean * 1e6}
        if cx.world == 1:
            r["us"] = s_mean * 1e6
        else:
            modes = {}
            for halo in ("fused", "exchange", "recompute"):
                kw = dict(halo=halo, comm=comm if halo != "exchange" else None)
                cx.barrier()
                lo, hi, Tb = assemble_interval_sharded(B, B, 1, 1, cx.rank, cx.world, **kw)
                lo, hi, Vb = assemble_interval_sharded(B, B, 0, 0, cx.rank, cx.world, coulomb_Z=1.0, **kw)
                eT = relerr(Tb.cpu().numpy(), full_T[lo:hi].cpu().numpy())
                eV = relerr(Vb.cpu().numpy(), full_V[lo:hi].cpu().numpy())
                okw = dict(kw)
                if halo != "exchange":
                    okw["out"] = None

                def sharded(_s, Tb=Tb, Vb=Vb, kw=kw, halo=halo):
                    if halo == "exchange":
                        assemble_interval_sharded(B, B, 0, 0, cx.rank, cx.world, coulomb_Z=1.0, **kw)
                        assemble_interval_sharded(B, B, 1, 1, cx.rank, cx.world, **kw)
                    else:
                        assemble_interval_sharded(B, B, 0, 0, cx.rank, cx.world, coulomb_Z=1.0, out=Vb, **kw)
                        assemble_interval_sharded(B, B, 1, 1, cx.rank, cx.world, out=Tb, **kw)
                mean, mn = cx.dev_time(sharded, 10, warm=3)
                err = cx.maxr(max(eT, eV))
                stt = cx.maxr(comm.status())
                modes[halo] = {"us": mean * 1e6, "us_min": mn * 1e6, "max_rel_vs_single_gpu": err, "comm_status": int(stt),
                               "speedup_vs_single_gpu": s_mean / mean}
                assert err <= 1e-13 and stt == 0, (halo, err, stt)
            r["halo"] = modes
            best = min(modes, key=lambda h: modes[h]["us"])
            r["us"] = modes[best]["us"]
            r["best_halo"] = best
        sec = r["us"] * 1e-6
        r["fp64_tflops"] = C3_FLOP * scale / sec / 1e12
        r["roofline"] = cx.fp64_roofline(C3_FLOP * scale / cx.world, C3_BYTES * scale / cx.world, sec, "bspline_assemble_stream_kernel<10,0> + <10,1> (V + T)")
        if nint == 200_000 and cx.rank == 0:
            cbo = _cbo(True)
            te = cbo.expand_knots(np.asarray(B.t.t), K)
            xq, wq = cbo.lgwt(te, 11)
            Vo, _, _ = cbo.bspline_assemble(te, K, te, K, xq, wq, 11, 0, 0, -1.0 / xq)
            To, _, _ = cbo.bspline_assemble(te, K, te, K, xq, wq, 11, 1, 1)
            r["parity"] = {"vs": "oracle bspline_assemble (every band entry of the timed output)", "V_max_rel": relerr(full_V.cpu().numpy(), Vo),
                           "T_max_rel": relerr(full_T.cpu().numpy(), To)}
            assert max(r["parity"]["V_max_rel"], r["parity"]["T_max_rel"]) <= RTOL, r["parity"]
            # e2e: the two bands into host BandedMatrix arrays (cb_bspline_assemble_host), pinned destinations
            hV, hT = pinned(torch, (B.nf, W)), pinned(torch, (B.nf, W))

            def e2e_c3():
                _lib.call("cb_bspline_assemble_host", B.handle, B.handle, 0, 0, None, 1, 1.0, 0, B.nf, 0, B.nf, K - 1, K - 1, hptr(hV), st)
                _lib.call("cb_bspline_assemble_host", B.handle, B.handle, 1, 1, None, 0, 0.0, 0, B.nf, 0, B.nf, K - 1, K - 1, hptr(hT), st)
            e2e_c3()
            t0 = time.perf_counter()
            for _ in range(5):
                e2e_c3()
            e2e_s = (time.perf_counter() - t0) / 5
            r["e2e"] = {"value": C3_FLOP / e2e_s / 1e12, "unit": "TFLOP/s", "ms": e2e_s * 1e3, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 2 * 8 * B.nf * W, "entry": "cb_bspline_assemble_host x2 (pinned host bands), rank 0 alone",
                        "max_rel_vs_oracle": max(relerr(hV.numpy(), Vo), relerr(hT.numpy(), To))}
            assert r["e2e"]["max_rel_vs_oracle"] <= RTOL
            del hV, hT
            if cpu:
                Bs_t = np.asarray(cb.ExpKnotSet(K, -3.0, 2.0, 20_000).t)
                tes = cbo.expand_knots(Bs_t, K)
                xs, ws = cbo.lgwt(tes, 11)
                cbo.set_threads(1)
                s1, _ = _best(lambda: (cbo.bspline_assemble(tes, K, tes, K, xs, ws, 11, 0, 0, -1.0 / xs),
                                       cbo.bspline_assemble(tes, K, tes, K, xs, ws, 11, 1, 1)), 2)
                cbo.set_threads(cbo.max_threads())
                sN, _ = _best(lambda: (cbo.bspline_assemble(te, K, te, K, xq, wq, 11, 0, 0, -1.0 / xq),
                                       cbo.bspline_assemble(te, K, te, K, xq, wq, 11, 1, 1)), 2)
                tl = np.asarray(cb.ExpKnotSet(K, -3.0, 2.0, 2_000).t)
                tel = cbo.expand_knots(tl, K)
                xl, wl = cbo.lgwt(tel, 11)
                sl, _ = _best(lambda: (cbo.bspline_assemble_literal(tel, K, tel, K, xl, wl, 11, 0, 0, -1.0 / xl),
                                       cbo.bspline_assemble_literal(tel, K, tel, K, xl, wl, 11, 1, 1)), 1)
                r["cpu_baseline"] = {"value": C3_FLOP / sN / 1e12, "unit": "TFLOP/s", "ms": sN * 1e3, "cores": int(cbo.max_threads()), "kind": "port",
                                     "sample": "V + T at the full 2e5 intervals, linear-cost restatement, OpenMP over intervals, best of 2",
                                     "single_thread_value": C3_FLOP * 0.1 / s1 / 1e12, "single_thread_ms_scaled_to_full": s1 * 10 * 1e3,
                                     "single_thread_sample": "V + T on ExpKnotSet(10,-3,2,20000) (1/10 of the intervals), best of 2",
                                     "literal_algorithm": {"ms": sl * 1e3, "intervals": 2000, "cores": 1,
                                                           "sample": "the reference's sparse basis_functions + overlap_matrix! "
                                                                     "(src/bsplines.jl:28-93) at 2e3 intervals, V + T"}}
        sizes[tag] = r
        del B, full_V, full_T
    out.update({"metric": "band_assembly_fp64_TFLOPs", "unit": "TFLOP/s", "value": sizes["2e5"]["fp64_tflops"], "us": sizes["2e5"]["us"],
                "roofline": sizes["2e5"]["roofline"], "parity": sizes["2e5"].get("parity"), "e2e": sizes["2e5"].get("e2e"),
                "cpu_baseline": sizes["2e5"].get("cpu_baseline"), "sizes": sizes})
    return out


# --------------------------------------------------------------------------------------
# C4: StaggeredFiniteDifferences N = 1e7, tridiagonal Laplacian applied to 4096 vectors (vectors split over the ranks)
# --------------------------------------------------------------------------------------
def config_c4(cx: Ctx, cpu: bool):
    torch, cb = cx.torch, cx.cb
    from compactbases_jl_b200.sharding import batch_range
    N, NV, TILE = 10_000_000, 4096, 256
    R = cb.StaggeredFiniteDifferences(N, 1e-3)
    D = cb.Derivative()
    Lp = R.T @ D.T @ D @ R
    lo, hi = batch_range(NV, cx.rank, cx.world)
    ntile = -(-(hi - lo) // TILE)
    X = torch.empty((TILE, N), dtype=torch.float64, device="cuda")
    Y = torch.empty_like(X)
    par = {}
    cbo = _cbo(True) if cx.rank == 0 else None
    if cx.rank == 0:
        al, be, _, _ = cbo.fd_staggered_uniform(N)
        dv, ev = cbo.fd_laplacian(al, be, 1e-3)
        par["stencils_bit_exact"] = bool(np.array_equal(Lp.dv.cpu().numpy(), dv) and np.array_equal(Lp.ev.cpu().numpy(), ev))
        par["max_rel"] = 0.0
        par["vectors_checked"] = []
    evs = []
    # the 4096 x 1e7 batch (328 GB) exceeds HBM: it is produced tile by tile by a counter-based generator (Philox, seed 4 +
    # global tile index) and every tile is applied once; the events bracket the apply launches only
    cx.barrier()
    for it in range(ntile):
        v0 = lo + it * TILE
        nv = min(TILE, hi - v0)
        g = torch.Generator(device="cuda").manual_seed(4 + v0 // TILE)
        X[:nv].normal_(generator=g)
        if it == 0:
            cb.mul(Y[:nv], Lp, X[:nv])                       # warm-up on the first tile
        # every tile is applied twice and the shorter device interval counts: an event pair also measures the time the
        # GPU waits for the HOST when its queue runs empty (a descheduled Python thread once showed up as a 365 ms "kernel")
        pair = []
        for _rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); cb.mul(Y[:nv], Lp, X[:nv]); e1.record()
            pair.append((e0, e1))
        evs.append(pair)
        if cx.rank == 0 and it in (0, ntile - 1):
            for v in (0, nv - 1):
                ref = cbo.tridiag_apply(ev, dv, ev, X[v].cpu().numpy()[:, None])[:, 0]
                par["max_rel"] = max(par["max_rel"], relerr(Y[v].cpu().numpy(), ref))
                par["vectors_checked"].append(int(v0 + v))
    torch.cuda.synchronize()
    tile_ms = [min(a.elapsed_time(b) for a, b in pair) for pair in evs]
    sec = cx.maxr(sum(tile_ms) * 1e-3)
    byt_rank = 2 * 8 * N * (hi - lo) + ntile * 2 * 8 * N
    byt = 2 * 8 * N * NV
    out = {"workload": "C4: StaggeredFiniteDifferences N=1e7 (rho=1e-3) tridiagonal Laplacian applied to 4096 vectors (655 GB of X + Y)",
           "scaling": "strong", "sharding": "vector batch split over the ranks, diagonals replicated, no collective",
           "metric": "tridiag_apply_algorithmic_GBps", "unit": "GB/s", "value": byt / sec / 1e9, "ms": sec * 1e3,
           "tiles_per_rank": ntile, "tile_vectors": TILE,
           "tile_ms_rank0": [round(float(x), 3) for x in tile_ms],
           "l2": "every 256-vector tile (20.5 GB in, 20.5 GB out) is larger than L2 and is regenerated before its launches (two per tile, the shorter one counts)",
           "roofline": cx.hbm_roofline(byt_rank / ntile, sec / ntile, "tridiag_apply_pipe_kernel", "per rank and tile: X + Y + the two diagonals")}
    if cx.rank == 0:
        par["vs"] = "oracle tridiag_apply on whole vectors of the first and last timed tile"
        assert par["stencils_bit_exact"] and par["max_rel"] <= RTOL, par
        out["parity"] = par
    del X, Y
    torch.cuda.empty_cache()
    # ---- e2e: host X -> device -> host Y (cb_tridiag_apply_host), 32 vectors per rank (2.56 GB each way) ----
    nve = 32
    Xh, Yh = pinned(torch, (nve, N)), pinned(torch, (nve, N))
    Xh.normal_()
    Xn, Yn = Xh.numpy().T, Yh.numpy().T
    sec_e = cx.wall_time(lambda: Lp.mul_host(Yn, Xn), 2, warm=1)
    out["e2e"] = {"value": 2 * 8 * N * nve * cx.world / sec_e / 1e9, "unit": "GB/s", "ms": sec_e * 1e3, "h2d_bytes_per_step": 8 * N * nve,
                  "d2h_bytes_per_step": 8 * N * nve, "entry": "cb_tridiag_apply_host (pinned host X, Y)",
                  "sample": f"{nve} vectors per rank (the full batch is 328 GB of host memory per direction)"}
    if cx.rank == 0:
        ref = cbo.tridiag_apply(ev, dv, ev, Xn[:, nve - 1:nve].copy(order="F"))[:, 0]
        out["e2e"]["max_rel_vs_oracle"] = relerr(Yn[:, nve - 1], ref)
        assert out["e2e"]["max_rel_vs_oracle"] <= RTOL
    del Xh, Yh
    if cpu and cx.rank == 0:
        Xc = np.asfortranarray(np.random.default_rng(4).standard_normal((N, 64)))
        Yc = np.zeros_like(Xc)
        cbo.set_threads(1)
        s1, _ = _best(lambda: cbo.tridiag_apply(ev, dv, ev, Xc[:, :8], 1.0, 0.0, Yc[:, :8]), 3)
        cbo.set_threads(cbo.max_threads())
        sN, _ = _best(lambda: cbo.tridiag_apply(ev, dv, ev, Xc, 1.0, 0.0, Yc, threads=True), 3)
        out["cpu_baseline"] = {"value": 2 * 8 * N * 64 / sN / 1e9, "unit": "GB/s", "cores": int(cbo.max_threads()), "kind": "port",
                               "sample": "64 of the 4096 vectors, OpenMP over vectors, best of 3",
                               "single_thread_value": 2 * 8 * N * 8 / s1 / 1e9, "single_thread_sample": "8 vectors, best of 3",
                               "literal_algorithm": "the stdlib SymTridiagonal mul! is already linear: same loop"}
    return out


# --------------------------------------------------------------------------------------
# C5: mutual densities of 1e4 orbital pairs on BSpline k = 8, 5e4 intervals (pairs split over the ranks)
# --------------------------------------------------------------------------------------
C5_FLOP = 2.3e11


def config_c5(cx: Ctx, cpu: bool):
    torch, cb = cx.torch, cx.cb
    from compactbases_jl_b200.sharding import batch_range
    K, NINT, P = 8, 50_000, 10_000
    t = cb.LinearKnotSet(K, 0, 500, NINT)
    B = cb.BSpline(t)
    nf = B.nf
    lo, hi = batch_range(P, cx.rank, cx.world)
    p = hi - lo
    rho = cb.Density(B, nlhs=p, nrhs=p)
    env = torch.exp(-((torch.linspace(0, 1, nf, dtype=torch.float64, device="cuda") - 0.5) / 0.25) ** 2)
    g5 = torch.Generator(device="cuda").manual_seed(5 + cx.rank)
    f = torch.randn((p, nf), dtype=torch.float64, device="cuda", generator=g5) * env
    g = torch.randn((p, nf), dtype=torch.float64, device="cuda", generator=g5) * env
    o = cb.ColMajor.zeros(nf, p)
    fc, gc = cb.ColMajor(f), cb.ColMajor(g)
    mean, mn = cx.dev_time(lambda s: rho.copyto(fc, gc, o), 3, warm=1)
    byt = 3 * 8 * nf * P
    out = {"workload": "C5: Density (mutual-density projection) of 1e4 orbital pairs, BSpline k=8, LinearKnotSet(8,0,500,50000) "
                       "(nf=50 007, q=9)", "scaling": "strong", "sharding": "pair batch split over the ranks, plan replicated, no collective",
           "metric": "density_pairs_per_s", "unit": "pairs/s", "value": P / mean, "ms": mean * 1e3, "ms_min": mn * 1e3,
           "fp64_tflops": C5_FLOP / mean / 1e12, "GBps": byt / mean / 1e9,
           "l2": "f, g, rho of one rank are 3 x %.1f GB, larger than L2" % (8 * nf * p / 1e9),
           "roofline": cx.fp64_roofline(C5_FLOP * p / P, byt * p / P, mean, "density_rhs_kernel<8,0,9> + partitioned banded solve (6 launches)")}
    if cx.rank == 0:
        cbo = _cbo(True)
        te = cbo.expand_knots(t.t, K)
        xq, wq = cbo.lgwt(te, 9)
        sel = [0, 1, p // 2, p - 1]
        ref = cbo.density_bspline_qr(te, K, xq, 9, np.asfortranarray(f[sel].cpu().numpy().T), np.asfortranarray(g[sel].cpu().numpy().T))
        got = o.data[sel].cpu().numpy().T
        out["parity"] = {"vs": "oracle density_bspline_qr (banded least squares) on pairs %s of the timed output" % sel,
                         "max_rel": max(relerr(got[:, i], ref[:, i]) for i in range(len(sel)))}
        assert out["parity"]["max_rel"] <= RTOL, out["parity"]
    # ---- e2e: host f, g -> device -> host rho (cb_density_apply_host), 1024 pairs in total ----
    pe = max(32, 1024 // cx.world)
    fh, gh, rh = pinned(torch, (pe, nf)), pinned(torch, (pe, nf)), pinned(torch, (pe, nf))
    fh.copy_(f[:pe]); gh.copy_(g[:pe])
    fn_, gn_, rn_ = fh.numpy().T, gh.numpy().T, rh.numpy().T
    sec_e = cx.wall_time(lambda: rho.copyto_host(fn_, gn_, rn_), 3, warm=1)
    out["e2e"] = {"value": pe * cx.world / sec_e, "unit": "pairs/s", "ms": sec_e * 1e3, "h2d_bytes_per_step": 2 * 8 * nf * pe,
                  "d2h_bytes_per_step": 8 * nf * pe, "entry": "cb_density_apply_host (pinned host f, g, rho)", "sample": f"{pe} pairs per rank",
                  "max_abs_diff_vs_device_path": float((rh.cuda() - o.data[:pe]).abs().max())}
    if cpu and cx.rank == 0:
        fc_ = np.asfortranarray(f[:64].cpu().numpy().T); gc_ = np.asfortranarray(g[:64].cpu().numpy().T)
        cbo.set_threads(1)
        s1, _ = _best(lambda: cbo.density_bspline_qr(te, K, xq, 9, fc_[:, :8], gc_[:, :8]), 2)
        cbo.set_threads(cbo.max_threads())
        sN, _ = _best(lambda: cbo.density_bspline_qr(te, K, xq, 9, fc_, gc_), 2)
        tl = cbo.linear_knots(0, 2, 200)
        tel = cbo.expand_knots(tl, K)
        xl, wl = cbo.lgwt(tel, 9)
        Vl = cbo.basis_matrix(tel, K, xl, 9)
        fl = np.random.default_rng(5).standard_normal((Vl.shape[1], 16)); gl = np.random.default_rng(6).standard_normal((Vl.shape[1], 16))
        sl, _ = _best(lambda: cbo.density_bspline(Vl, fl, gl), 1)
        out["cpu_baseline"] = {"value": 64 / sN, "unit": "pairs/s", "cores": int(cbo.max_threads()), "kind": "port",
                               "sample": "64 of the 1e4 pairs, linear-cost banded least squares, OpenMP, best of 2",
                               "single_thread_value": 8 / s1, "single_thread_sample": "8 pairs, best of 2",
                               "literal_algorithm": {"value": 16 / sl, "unit": "pairs/s", "cores": 1,
                                                     "sample": "the reference's dense pinv(Matrix(RV)) (src/densities.jl:96) at 200 intervals x 16 pairs"}}
    del f, g, o, fh, gh, rh
    return out


def other_kernels(cx: Ctx):
    """Short device-resident timings of kernels around the configs (rank 0): banded apply, metric solves."""
    torch, cb = cx.torch, cx.cb
    D = cb.Derivative()
    peak = cx.peak_gbs
    out = {}
    B = cb.BSpline(cb.LinearKnotSet(7, 0, 100, 1000))
    S1 = B.T @ D.T @ D @ B
    nv = 65536
    X = torch.randn((nv, B.nf), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    mean, mn = cx_dev_time_local(torch, lambda: cb.mul(Y, S1, X), 5)
    byt = 2 * 8 * B.nf * nv
    out["c1_banded_apply_65536vec"] = {"us": mean * 1e6, "GBps": byt / mean / 1e9, "hbm_frac": byt / mean / 1e9 / peak}
    cbo = _cbo(True)
    ref = cbo.banded_apply(S1.data.cpu().numpy(), B.nf, B.nf, 6, 6, np.asfortranarray(X[-2:].cpu().numpy().T))
    out["c1_banded_apply_65536vec"]["parity_max_rel"] = relerr(Y[-2:].cpu().numpy().T, ref)
    assert out["c1_banded_apply_65536vec"]["parity_max_rel"] <= RTOL
    del X, Y
    B3 = cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, 200_000))
    T3 = B3.T @ D.T @ D @ B3
    nv = 256
    X = torch.randn((nv, B3.nf), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    mean, mn = cx_dev_time_local(torch, lambda: cb.mul(Y, T3, X), 5)
    byt = 2 * 8 * B3.nf * nv + 8 * 19 * B3.nf
    out["c3_banded_apply_256vec"] = {"us": mean * 1e6, "GBps": byt / mean / 1e9, "hbm_frac": byt / mean / 1e9 / peak}
    ref = cbo.banded_apply(T3.data.cpu().numpy(), B3.nf, B3.nf, 9, 9, np.asfortranarray(X[-2:].cpu().numpy().T))
    out["c3_banded_apply_256vec"]["parity_max_rel"] = relerr(Y[-2:].cpu().numpy().T, ref)
    assert out["c3_banded_apply_256vec"]["parity_max_rel"] <= RTOL
    del X, Y, B3, T3
    B5 = cb.BSpline(cb.LinearKnotSet(8, 0, 500, 50_000))
    P = 1024
    S5 = B5.T @ B5
    Xs = torch.randn((P, B5.nf), dtype=torch.float64, device="cuda")
    byt = 2 * 8 * B5.nf * P
    for tag, F in (("c5_bandchol_solve_1024vec", S5.cholesky()), ("c5_bandlu_solve_1024vec", cb.BandLU(S5))):
        mean, mn = cx_dev_time_local(torch, lambda: F.ldiv(Xs), 3, warm=1)
        out[tag] = {"ms": mean * 1e3, "GBps": byt / mean / 1e9, "hbm_frac": byt / mean / 1e9 / peak}
        Xs.normal_()
    return out


def fp64_peak_tflops():
    try:
        lib = C.CDLL(os.path.join(ROOT, "compactbases_jl_b200", "libcb_micro.so"))
        v = C.c_double(0.0)
        if lib.cbm_fp64_peak(C.byref(v)) == 0 and v.value > 1.0:
            return float(v.value)
    except OSError:
        pass
    return 37.0             # nominal B200 FP64 if the helper is missing


def run_gpu(args):
    # exactly one line may reach stdout (the JSON): libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # pinned staging buffers next to the GPU: run on (and first-touch from) the CPUs of the GPU's NUMA node
    from compactbases_jl_b200.numa import bind_to_gpu_numa
    global _ALL_CPUS
    _ALL_CPUS = os.sched_getaffinity(0)
    numa_info = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import compactbases_jl_b200 as cb
    from compactbases_jl_b200 import _lib
    _lib.load()                                             # fails loudly if the CUDA library is missing
    peaks, peak_src = measured_peaks()
    cx = Ctx(torch, dist, cb, rank, world, peaks, peak_src, fp64_peak_tflops())
    peak_gbs = cx.peak_gbs
    want = set(args.configs.split(",")) if args.configs else {"c1", "c3", "c4", "c5", "kernels"}
    if args.no_extras:
        want = set()

    F = cb.FEDVR(np.linspace(0.0, 1000.0, NEL + 1), ORDER)
    assert F.N == N_FUN
    gen = torch.Generator(device="cuda").manual_seed(2 + rank)
    X = torch.randn((NVEC, N_FUN), dtype=torch.float64, device="cuda", generator=gen)
    Y = torch.empty_like(X)
    Xc, Yc = cb.ColMajor(X), cb.ColMajor(Y)
    st = torch.cuda.current_stream()

    def step():
        return F.derop_mul(2, Yc, Xc)                       # K4 + K8 in one foreign call: re-assemble, then apply

    for _ in range(max(args.warmup, 3)):
        A = step()
    # apply kernel alone (roofline): events around cb_fedvr_apply launches on the operator just assembled
    apply_s, _ = cx.dev_time(lambda s: cb.mul(Yc, A, Xc), max(5, min(args.steps, 20)), warm=1)
    sampler = ClockSampler(local) if rank == 0 else None
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    cx.barrier()
    elapsed = cx.maxr(e0.elapsed_time(e1) * 1e-3)
    value = BYTES_STEP * world * args.steps / elapsed / 1e9

    # K4 shards: every rank assembles its element range; bit-identical to its slice of the full assembly
    from compactbases_jl_b200.sharding import fedvr_assemble_sharded
    full_blocks = F.derop(2)
    e_lo, e_hi, part = fedvr_assemble_sharded(F, 2, rank, world)
    shard_ok = bool(torch.equal(part, full_blocks[e_lo * ORDER * ORDER:e_hi * ORDER * ORDER]))
    asm_full_s, _ = cx.dev_time(lambda s: _lib.call("cb_fedvr_assemble", F.d_x.data_ptr(), F.d_wcat.data_ptr(), F.d_n.data_ptr(), None, None,
                                                    None, None, NEL, ORDER, ORDER, 2, full_blocks.data_ptr(), torch.cuda.current_stream().cuda_stream), 20)
    asm_part_s, _ = cx.dev_time(lambda s: fedvr_assemble_sharded(F, 2, rank, world, out=part), 20)
    sharded_assembly = {"elements_per_rank": e_hi - e_lo, "us_full": asm_full_s * 1e6, "us_sharded": asm_part_s * 1e6,
                        "bit_identical": bool(cx.maxr(0.0 if shard_ok else 1.0) == 0.0)}
    assert sharded_assembly["bit_identical"]

    parity = None
    if rank == 0:                                           # the timed output against the oracle: first, middle and last vector
        cbo = _cbo(True)
        O = cbo.FEDVROracle(np.linspace(0.0, 1000.0, NEL + 1), ORDER)
        sel = [0, NVEC // 2, NVEC - 1]
        ref = O.apply(O.blocks(2), np.asfortranarray(X[sel].cpu().numpy().T))
        parity = {"vs": "oracle FEDVROracle.blocks + apply on vectors %s of the timed output" % sel,
                  "max_rel": relerr(Y[sel].cpu().numpy().T, ref),
                  "blocks_max_rel": relerr(A.blocks.cpu().numpy(), O.blocks(2))}
        assert max(parity["max_rel"], parity["blocks_max_rel"]) <= RTOL, parity

    # ---- end to end through the host-buffer C-ABI entry (pinned host memory) ----
    Xh = pinned(torch, (NVEC, N_FUN))
    Yh = pinned(torch, (NVEC, N_FUN))
    Xh.copy_(X)
    Xn, Yn = Xh.numpy().T, Yh.numpy().T                      # (n, nvec) Fortran-ordered views of the pinned buffers

    Dv = cb.Derivative()

    def e2e_step():
        F._ops.clear()
        Ae = F.T @ Dv.T @ Dv @ F
        Ae.mul_host(Yn, Xn)                                  # orders itself after the assembly on the current stream
        return float(Yn[0, 0])

    e2e_steps = max(3, min(args.steps, 10))
    e2e_s = cx.wall_time(e2e_step, e2e_steps, warm=2)
    e2e_value = BYTES_STEP * world / e2e_s / 1e9
    ref_chk = float(torch.max(torch.abs(Yh.cuda() - Y)).item())   # host path == device path
    del Xh, Yh
    clocks = sampler.stop() if sampler else None

    cpu = world == 1
    configs = {}
    comm = None
    if world > 1 and "c3" in want:
        from compactbases_jl_b200.sharding import Comm
        comm = Comm(rank, world)
    del X, Y, Xc, Yc
    torch.cuda.empty_cache()
    for name, fn in (("C1", lambda: config_c1(cx, cpu)), ("C3", lambda: config_c3(cx, cpu, comm)),
                     ("C4", lambda: config_c4(cx, cpu)), ("C5", lambda: config_c5(cx, cpu))):
        if name.lower() in want:
            cx.barrier()                                    # ranks enter every config together (rank 0 alone does the CPU legs)
            t0 = time.perf_counter()
            configs[name] = fn()
            configs[name]["n_gpus"] = world
            configs[name]["bench_wall_s"] = time.perf_counter() - t0
            torch.cuda.empty_cache()
    kernels = other_kernels(cx) if (rank == 0 and "kernels" in want) else {}
    if world > 1:
        cx.barrier()

    if rank == 0:
        cpu_v, cpu_sec, cores = cpu_reference_pass(128, repeats=2)
        cpu_1, _, _ = cpu_reference_pass(16, repeats=2, threads=False)
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tj = json.load(f)
                traffic = tj.get("fedvr_apply_kernel_dram_bytes_per_launch")
                traffic_src = tj.get("source")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_functions": N_FUN, "vectors_per_gpu": NVEC, "bytes_per_step_per_gpu": BYTES_STEP,
                       "l2": "inputs larger than L2 (X and Y are 737 MB each)", "sharding": "vector batch per rank, operator replicated, no collective",
                       "step": "cb_fedvr_assemble_apply (one foreign call: K4 assembly + K8 apply)",
                       "k4_element_shards": sharded_assembly},
            "roofline": {"bound": "hbm", "achieved": BYTES_APPLY / apply_s / 1e9, "peak": peak_gbs, "unit": "GB/s",
                         "frac": BYTES_APPLY / apply_s / 1e9 / peak_gbs, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "fedvr_apply_pipe_kernel<10,16>",
                         "algorithmic_bytes_per_launch": BYTES_APPLY, "kernel_ms": apply_s * 1e3, "peak_source": cx.peak_src},
            "cpu_baseline": {"value": cpu_v, "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": "assemble + apply on 128 of the 1024 vectors, best of 2 (oracle/, OpenMP)",
                             "single_thread_value": cpu_1, "single_thread_sample": "16 of the 1024 vectors, best of 2"},
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": 8 * N_FUN * NVEC, "d2h_bytes_per_step": 8 * N_FUN * NVEC,
                    "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "entry": "cb_fedvr_apply_host (pinned host X, Y)",
                    "max_abs_diff_vs_device_path": ref_chk},
            "parity": parity,
            "gpu_launches": 2 * args.steps,
            "clocks": clocks,
            "fp64_peak_tflops_measured": cx.fp64_peak,
            "host": {"cpus": os.cpu_count(), "numa_binding_rank0": numa_info},
            "configs": configs,
            "kernels": kernels,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (skip the other configs)")
    ap.add_argument("--configs", default="", help="comma-separated subset of c1,c3,c4,c5,kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
