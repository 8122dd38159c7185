"""Multi-GPU check (torchrun, one rank per GPU): interval-sharded assembly with the halo exchange fused into the
kernel (peer-mapped mailboxes, cb_bspline_assemble_sharded), with the NCCL send/recv baseline and with halo recompute
vs the single-GPU assembly; times all three at the C3 size and at 8x the work."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import compactbases_jl_b200 as cb
from compactbases_jl_b200.sharding import Comm, assemble_interval_sharded, interval_shards, allgather_columns
comm = Comm(rank, world)

def tmax(ms):
    t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

D = cb.Derivative()
for nint in (200_000, 1_600_000):
    B = cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, nint))
    full_T = (B.T @ D.T @ D @ B).data
    full_V = (B.T @ cb.CoulombPotential(1.0) @ B).data
    for halo in ("fused", "exchange", "recompute"):
        kw = dict(halo=halo, comm=comm if halo != "exchange" else None)
        lo, hi, T = assemble_interval_sharded(B, B, 1, 1, rank, world, **kw)
        lo, hi, V = assemble_interval_sharded(B, B, 0, 0, rank, world, coulomb_Z=1.0, **kw)
        eT = (T - full_T[lo:hi]).abs().max().item() / full_T.abs().max().item()
        eV = (V - full_V[lo:hi]).abs().max().item() / full_V.abs().max().item()
        assert eT <= 1e-13 and eV <= 1e-13, (halo, eT, eV)
        Tb, Vb = torch.empty_like(T), torch.empty_like(V)
        for _ in range(3):
            Tb = assemble_interval_sharded(B, B, 1, 1, rank, world, out=Tb, **kw)[2]
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            Tb = assemble_interval_sharded(B, B, 1, 1, rank, world, out=Tb, **kw)[2]     # (halo="exchange" returns a new tensor)
            Vb = assemble_interval_sharded(B, B, 0, 0, rank, world, coulomb_Z=1.0, out=Vb, **kw)[2]
        e1.record(); torch.cuda.synchronize()
        ms = tmax(e0.elapsed_time(e1) / reps)
        assert comm.status() == 0
        assert torch.equal(Tb, T) or (Tb - T).abs().max().item() <= 1e-13 * full_T.abs().max().item()
        if rank == 0:
            print(f"C3-like nint={nint} world={world} halo={halo}: V+T {ms*1e3:.1f} us (max over ranks), rel err {max(eT, eV):.1e}", flush=True)
    shards = interval_shards(B.t.numintervals(), 10, 10, 0, B.nf, world)
    G = allgather_columns(T, shards)
    assert (G - full_T).abs().max().item() <= 1e-13 * full_T.abs().max().item()
    del B, full_T, full_V
if rank == 0:
    print("mgpu check ok", flush=True)
dist.destroy_process_group()
