#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (contract: see DESIGN.md §Measurement).

Workload at N GPUs (BASELINE.json configs[1], "C2"): FE-DVR order 10 with 1e4 elements
(n = 90 001 functions); one STEP = assemble the block-banded second-derivative operator
B'D'DB (K4) and apply it to a batch of 1024 coefficient vectors (K8), all Float64.  The vector
batch is the sharded dimension: every rank holds its own 1024 vectors and its own copy of the
operator (no data-path collective, weak scaling).

  value   algorithmic GB/s of the whole job with inputs resident in HBM
          (bytes/step = read X + write Y + element blocks written once and read once)
  e2e     the same metric through the host-buffer C-ABI entry (cb_fedvr_apply_host): pinned
          host X -> device -> Y back to the host inside the timed region
  roofline  the apply kernel alone (CUDA events around its launches) vs the measured HBM peak
  cpu_baseline / --impl reference   the CPU restatement of the reference (oracle/, OpenMP over
          all host cores) on a bounded slice of the same workload

Other kernels of the path (B-spline evaluation, band assembly, tridiagonal / banded apply,
density projection) are timed briefly and reported under "kernels" so that every BASELINE config
has a measured number in the same JSON line.
"""
from __future__ import annotations

import argparse
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


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


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
# CPU reference arm: the oracle's restatement of the reference algorithm, all host threads
# --------------------------------------------------------------------------------------
_CPU_STATE = {}


def cpu_reference_pass(nvec_sample, repeats=3, threads=True):
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every host core
    if threads and "oracle.cbo" not in sys.modules:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import cbo
    if nvec_sample not in _CPU_STATE:                       # basis construction + inputs are outside the timed region
        t = np.linspace(0.0, 1000.0, NEL + 1)
        rng = np.random.default_rng(2)
        _CPU_STATE[nvec_sample] = (cbo.FEDVROracle(t, ORDER), np.asfortranarray(rng.standard_normal((N_FUN, nvec_sample))),
                                   np.zeros((N_FUN, nvec_sample), order="F"))
    O, X, Y = _CPU_STATE[nvec_sample]
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        blk = O.blocks(2)                                   # assemble (single-threaded, like the reference)
        O.apply(blk, X, 1.0, 0.0, Y, threads=threads)       # apply over the vector slice
        best = min(best, time.perf_counter() - t0)
    bytes_sample = 2 * 8 * N_FUN * nvec_sample + 8 * NEL * ORDER * ORDER + BYTES_ASSEMBLE
    cores = int(cbo.lib().cbo_max_threads()) if threads else 1
    return bytes_sample / best / 1e9, best, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nvs = 256
    vals = []
    for _ in range(args.warmup):
        cpu_reference_pass(nvs, repeats=1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, sec, cores = cpu_reference_pass(nvs, repeats=1)
        vals.append((v, sec))
    total = time.perf_counter() - t0
    bytes_sample = 2 * 8 * N_FUN * nvs + 8 * NEL * ORDER * ORDER + BYTES_ASSEMBLE
    value = bytes_sample * args.steps / total / 1e9
    sample = f"assemble + apply on {nvs} of the {NVEC} vectors per step (CPU restatement of the reference, OpenMP)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def ev_time(torch, fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.mean(ts)), float(np.min(ts))


def other_kernels(torch, cb, peak_gbs):
    """Short device-resident timings of the rest of the path (one line per BASELINE config)."""
    from compactbases_jl_b200 import _lib
    from compactbases_jl_b200.operators import _ptr, _stream
    D = cb.Derivative()
    out = {}
    # C1: B-spline evaluation, k = 7, 1000 intervals, 1e6 points (68 B/point) and 1.6e7 points (> L2)
    B = cb.BSpline(cb.LinearKnotSet(7, 0, 100, 1000))
    for nx, tag in ((1_000_000, "c1_eval_1e6"), (16_000_000, "c1_eval_1.6e7")):
        x = torch.rand(nx, dtype=torch.float64, device="cuda") * 100
        iv = torch.empty(nx, dtype=torch.int32, device="cuda")
        vals = torch.empty((nx, 7), dtype=torch.float64, device="cuda")
        mean, mn = ev_time(torch, lambda: _lib.call("cb_bspline_eval", B.handle, _ptr(x), nx, 0, _ptr(iv), _ptr(vals), _stream()), 10)
        out[tag] = {"us": mn * 1e6, "evals_per_s": 7 * nx / mn, "points_per_s": nx / mn, "GBps": 68 * nx / mn / 1e9,
                    "hbm_frac": 68 * nx / mn / 1e9 / peak_gbs}
        del x, iv, vals
    # B[x,:] as Julia's SparseMatrixCSC (Int64 colptr/rowval, zeros dropped): evaluation + radix sort by column
    nx = 1_000_000
    x = torch.rand(nx, dtype=torch.float64, device="cuda") * 100
    mean, mn = ev_time(torch, lambda: B.eval_csc(x), 5)
    out["c1_eval_csc_1e6"] = {"us": mn * 1e6, "points_per_s": nx / mn, "note": "includes allocation of the CSC arrays and the nnz read-back"}
    del x
    mean, mn = ev_time(torch, lambda: (B.T @ B, B.T @ D.T @ D @ B), 10)
    out["c1_assemble_mass_laplacian"] = {"us": mn * 1e6, "note": "launch-latency bound (0.22 MB)"}
    # banded apply at the C1 basis: k = 7 band (l = u = 6), 1006 x 65 536 vectors
    S1 = B.T @ D.T @ D @ B
    nv = 65536
    X = torch.randn((nv, B.nf), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    mean, mn = ev_time(torch, lambda: cb.mul(Y, S1, X), 5)
    byt = 2 * 8 * B.nf * nv
    out["c1_banded_apply_65536vec"] = {"us": mn * 1e6, "GBps": byt / mn / 1e9, "hbm_frac": byt / mn / 1e9 / peak_gbs}
    del X, Y
    # C3: k = 10 exponential knots, 2e5 intervals, Coulomb + Laplacian bands (62.4 MB, ~1.6 GFLOP)
    B3 = cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, 200_000))
    mean, mn = ev_time(torch, lambda: (B3.T @ cb.CoulombPotential(1.0) @ B3, B3.T @ D.T @ D @ B3), 5)
    out["c3_assemble_coulomb_laplacian"] = {"us": mn * 1e6, "GBps": 62.4e6 / mn / 1e9, "hbm_frac": 62.4e6 / mn / 1e9 / peak_gbs,
                                            "fp64_tflops": 1.6e9 / mn / 1e12}
    # the two assembly kernels alone (C ABI, preallocated band: no Python operator glue in the timed region)
    band3 = torch.empty((B3.nf, 19), dtype=torch.float64, device="cuda")
    def asm_pair():
        _lib.call("cb_bspline_assemble_coulomb", B3.handle, B3.handle, 1.0, 0, B3.nf, 0, B3.nf, 9, 9, _ptr(band3), _stream())
        _lib.call("cb_bspline_assemble", B3.handle, B3.handle, 1, 1, None, 0, B3.nf, 0, B3.nf, 9, 9, _ptr(band3), _stream())
    mean, mn = ev_time(torch, asm_pair, 5)
    out["c3_assemble_coulomb_laplacian_cabi"] = {"us": mn * 1e6, "GBps": 62.4e6 / mn / 1e9, "hbm_frac": 62.4e6 / mn / 1e9 / peak_gbs,
                                                 "fp64_tflops": 1.6e9 / mn / 1e12}
    del band3
    T3 = B3.T @ D.T @ D @ B3
    nv = 256
    X = torch.randn((nv, B3.nf), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    mean, mn = ev_time(torch, lambda: cb.mul(Y, T3, X), 5)
    byt = 2 * 8 * B3.nf * nv + 8 * 19 * B3.nf
    out["c3_banded_apply_256vec"] = {"us": mn * 1e6, "GBps": byt / mn / 1e9, "hbm_frac": byt / mn / 1e9 / peak_gbs}
    del X, Y, B3, T3
    # C4: staggered FD tridiagonal Laplacian, N = 1e7, a 256-vector tile of the 4096 (41 GB of the 655 GB job)
    N, nv = 10_000_000, 256
    R = cb.StaggeredFiniteDifferences(N, 1e-3)
    Lp = R.T @ D.T @ D @ R
    X = torch.randn((nv, N), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    mean, mn = ev_time(torch, lambda: cb.mul(Y, Lp, X), 3, warm=1)
    byt = 2 * 8 * N * nv + 2 * 8 * N
    out["c4_tridiag_apply_256vec_tile"] = {"ms": mn * 1e3, "GBps": byt / mn / 1e9, "hbm_frac": byt / mn / 1e9 / peak_gbs,
                                           "full_job_ms_extrapolated": mn * 1e3 * 16}
    del X, Y, R, Lp
    # C5: density projection, k = 8, 5e4 intervals, 1024-pair tile of the 1e4 pairs
    B5 = cb.BSpline(cb.LinearKnotSet(8, 0, 500, 50_000))
    P = 1024
    rho = cb.Density(B5, nlhs=P, nrhs=P)
    f = torch.randn((P, B5.nf), dtype=torch.float64, device="cuda"); g = torch.randn((P, B5.nf), dtype=torch.float64, device="cuda")
    o = cb.ColMajor.zeros(B5.nf, P)
    mean, mn = ev_time(torch, lambda: rho.copyto(cb.ColMajor(f), cb.ColMajor(g), o), 3, warm=1)
    byt = 3 * 8 * B5.nf * P
    out["c5_density_1024pair_tile"] = {"ms": mn * 1e3, "GBps": byt / mn / 1e9, "hbm_frac": byt / mn / 1e9 / peak_gbs,
                                       "fp64_tflops": 2.3e11 * P / 1e4 / mn / 1e12}
    # metric inverse / shift-and-invert solves on the C5 basis: banded Cholesky (partitioned) and pivoted band LU,
    # 50 007 rows x 1024 vectors in place (410 MB read + written once, algorithmically)
    S5 = B5.T @ B5
    Xs = torch.randn((P, B5.nf), dtype=torch.float64, device="cuda")
    byt = 2 * 8 * B5.nf * P
    for tag, F in (("c5_bandchol_solve_1024vec", S5.cholesky()), ("c5_bandlu_solve_1024vec", cb.BandLU(S5))):
        mean, mn = ev_time(torch, lambda: F.ldiv(Xs), 3, warm=1)
        out[tag] = {"ms": mn * 1e3, "GBps": byt / mn / 1e9, "hbm_frac": byt / mn / 1e9 / peak_gbs}
        Xs.normal_()
    return out


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import compactbases_jl_b200 as cb
    from compactbases_jl_b200 import _lib
    _lib.load()                                             # fails loudly if the CUDA library is missing
    peaks, peak_src = measured_peaks()
    peak_gbs = float(peaks["hbm_gbs"])

    D = cb.Derivative()
    F = cb.FEDVR(np.linspace(0.0, 1000.0, NEL + 1), ORDER)
    assert F.N == N_FUN
    gen = torch.Generator(device="cuda").manual_seed(2 + rank)
    X = torch.randn((NVEC, N_FUN), dtype=torch.float64, device="cuda", generator=gen)
    Y = torch.empty_like(X)
    Xc, Yc = cb.ColMajor(X), cb.ColMajor(Y)
    apply_events = []

    def step(record=False):
        F._ops.clear()                                      # re-assemble every step (K4) ...
        A = F.T @ D.T @ D @ F
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        cb.mul(Yc, A, Xc)                                   # ... and apply (K8)
        if record:
            e1.record()
            apply_events.append((e0, e1))
        return A

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(record=True)
    e1.record()
    barrier()
    elapsed = e0.elapsed_time(e1) * 1e-3
    t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    apply_s = float(np.mean([a.elapsed_time(b) for a, b in apply_events])) * 1e-3
    value = BYTES_STEP * world * args.steps / elapsed / 1e9

    # ---- end to end through the host-buffer C-ABI entry (pinned host memory) ----
    A = step()
    Xh = torch.empty((NVEC, N_FUN), dtype=torch.float64, pin_memory=True)
    Yh = torch.empty((NVEC, N_FUN), dtype=torch.float64, pin_memory=True)
    Xh.copy_(X)
    Xn, Yn = Xh.numpy().T, Yh.numpy().T                      # (n, nvec) Fortran-ordered views of the pinned buffers

    def e2e_step():
        F._ops.clear()
        Ae = F.T @ D.T @ D @ F
        torch.cuda.current_stream().synchronize()
        Ae.mul_host(Yn, Xn)
        return float(Yn[0, 0])

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    t = torch.tensor([e2e_elapsed], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_elapsed = float(t.item())
    e2e_value = BYTES_STEP * world * e2e_steps / e2e_elapsed / 1e9
    ref_chk = float(torch.max(torch.abs(Yh.cuda() - Y)).item())   # host path == device path

    clocks = sampler.stop() if sampler else None
    if rank == 0:
        kernels = other_kernels(torch, cb, peak_gbs) if not args.no_extras else {}
        cpu_v, cpu_sec, cores = cpu_reference_pass(128, repeats=2)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get("fedvr_apply_kernel_dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_functions": N_FUN, "vectors_per_gpu": NVEC, "bytes_per_step_per_gpu": BYTES_STEP,
                       "l2": "inputs larger than L2 (X and Y are 737 MB each)", "sharding": "vector batch per rank, operator replicated, no collective"},
            "roofline": {"bound": "hbm", "achieved": BYTES_APPLY / apply_s / 1e9, "peak": peak_gbs, "unit": "GB/s",
                         "frac": BYTES_APPLY / apply_s / 1e9 / peak_gbs, "traffic": traffic, "kernel": "fedvr_apply_pipe_kernel<10,16>",
                         "algorithmic_bytes_per_launch": BYTES_APPLY, "kernel_ms": apply_s * 1e3, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)"},
            "cpu_baseline": {"value": cpu_v, "unit": "GB/s", "cores": cores, "kind": "port",
                             "sample": "assemble + apply on 128 of the 1024 vectors, best of 2 (oracle/, OpenMP)"},
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": 8 * N_FUN * NVEC, "d2h_bytes_per_step": 8 * N_FUN * NVEC,
                    "ms_per_step": e2e_elapsed / e2e_steps * 1e3, "steps": e2e_steps, "entry": "cb_fedvr_apply_host (pinned host X, Y)",
                    "max_abs_diff_vs_device_path": ref_chk},
            "gpu_launches": 2 * args.steps,
            "clocks": clocks,
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
    ap.add_argument("--no-extras", action="store_true", help="skip the short timings of the other kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
