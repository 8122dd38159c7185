"""World-size-2/3 gloo tests (CPU) of the multi-GPU plumbing: batch partitioning, interval
ownership and the halo exchange of an interval-sharded band assembly.  The per-rank partial
sums come from the oracle (quadrature weights zeroed outside the rank's interval range), so the
exchange is checked against the full oracle band without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cbo
from compactbases_jl_b200.sharding import batch_range, halo_exchange_add, interval_shards


def test_batch_range_partitions():
    for n in (0, 1, 7, 1024, 10_000):
        for world in (1, 2, 3, 8):
            r = [batch_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_interval_shards_cover_columns():
    for k, ml, nint, world in [(7, 7, 1000, 8), (10, 10, 200001, 8), (3, 1, 5, 2), (4, 2, 40, 3), (8, 8, 50000, 4)]:
        nf = nint + 1 + ml + k - 2 - k if False else None
        # nt = n_tt + ml + mr - 2 with mr = k;  nf = nt - k
        nf = (nint + 1) + ml + k - 2 - k
        sh = interval_shards(nint, k, ml, 0, nf, world)
        assert sh[0].own_lo == 0 and sh[-1].own_hi == nf
        for a, b in zip(sh[:-1], sh[1:]):
            assert a.own_hi == b.own_lo and a.s_hi == b.s_lo
        for s in sh:
            assert s.touch_lo <= s.own_lo and s.touch_hi == s.own_hi
            assert s.own_lo - s.touch_lo <= k - 1          # the halo is at most k-1 columns


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k, tt, ml, mr, a, b, col0, ncols_cut = case
        t = cbo.expand_knots(tt, k, ml, mr)
        q = cbo.num_quadrature_points(k, 3)
        x, w = cbo.lgwt(t, q)
        nf = len(t) - k
        ncols = nf - col0 - ncols_cut
        nint = len(tt) - 1
        shards = interval_shards(nint, k, ml, col0, ncols, world)
        sh = shards[rank]
        # partial sums of this rank: weights of the nodes outside its intervals set to zero
        nei = cbo.nonempty_intervals(t)                      # expanded 1-based indices of the non-empty intervals
        wm = w.copy()
        for c, I in enumerate(nei):
            s = I - ml                                       # t.t interval (0-based)
            if not (sh.s_lo <= s < sh.s_hi):
                wm[c * q:(c + 1) * q] = 0.0
        band, l, u = cbo.bspline_assemble(t, k, t, k, x, wm, q, a, b, rowL0=col0, m=ncols, colR0=col0, n=ncols)
        partial = torch.from_numpy(band[sh.touch_lo:sh.touch_hi].copy())
        # columns this rank does not touch must be zero in its partial band
        assert np.all(band[:sh.touch_lo] == 0) and np.all(band[sh.touch_hi:] == 0)
        own = halo_exchange_add(partial, sh, shards, rank)
        full, _, _ = cbo.bspline_assemble(t, k, t, k, x, w, q, a, b, rowL0=col0, m=ncols, colR0=col0, n=ncols)
        ref = full[sh.own_lo:sh.own_hi]
        err = float(np.max(np.abs(own.numpy() - ref))) if ref.size else 0.0
        out[rank] = err / max(float(np.max(np.abs(full))), 1e-300)
    finally:
        dist.destroy_process_group()


CASES = [
    (7, cbo.linear_knots(0, 10, 40), 7, 7, 1, 1, 0, 0),
    (4, cbo.exp_knots(-2.0, 1.0, 23), 4, 4, 0, 0, 1, 1),                    # restricted basis
    (3, np.array([0.0, 1, 1, 3, 4, 6, 7, 7.5, 9]), 1, 3, 0, 1, 0, 0),        # repeated knot, ml < k
]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_halo_exchange_matches_full_assembly(world, case):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, CASES[case], out), nprocs=world, join=True)
    assert len(out) == world
    assert max(out.values()) <= 1e-14
