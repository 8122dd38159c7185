"""Multi-GPU partitioning of the hot path (one process per GPU, ``torch.distributed``).

The reference is single-process; the partitioning follows SURVEY §8e:

* point batches (``B[x,:]``), coefficient-vector batches (``mul!``) and pair batches (``Density``)
  are split into contiguous ranges with **no exchange** (``batch_range``);
* quadrature assembly of a B-spline band can be split by **knot interval**.  A rank sums the
  contributions of its intervals into every column that touches them; a column (basis function)
  lives on k consecutive intervals, so the last k-1 columns a rank *owns* (those that start in
  its range) also receive contributions from the next rank's intervals.  Either each rank
  recomputes its own columns over the full interval range (``halo="recompute"``, no
  communication) or the neighbours add their partial sums: one send to the left and one
  receive from the right of at most k-1 columns x (l+u+1) rows (``halo="exchange"``; NCCL
  point-to-point over NVLink on GPUs, gloo in the CPU tests).

Only index arithmetic and the exchange live here; the kernels are behind
``cb_bspline_assemble_part`` in the C ABI.
"""
from __future__ import annotations

import ctypes as C
import functools
from dataclasses import dataclass

import torch
import torch.distributed as dist


def batch_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous share [lo, hi) of n independent items (points, vectors, pairs)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class IntervalShard:
    """What one rank computes in an interval-sharded assembly (0-based, half-open ranges)."""
    s_lo: int       # t.t intervals summed by this rank
    s_hi: int
    own_lo: int     # columns (within the restriction) whose first interval lies in [s_lo, s_hi): complete after the exchange
    own_hi: int
    touch_lo: int   # columns with any support in [s_lo, s_hi): computed as partial sums
    touch_hi: int


@functools.lru_cache(maxsize=256)
def interval_shards(numintervals: int, k: int, ml: int, col0: int, ncols: int, world: int) -> list[IntervalShard]:
    """Splits the t.t intervals evenly.  Function f (1-based, order k) lives on intervals
    f-ml .. f-ml+k-1 clipped to 0..numintervals-1 (src/knot_sets.jl:138-151), so it *starts* in
    interval max(0, f-ml)."""
    out = []
    for r in range(world):
        s_lo, s_hi = batch_range(numintervals, r, world)

        def start(j):            # first interval of column j (0-based within the restriction)
            return max(0, (col0 + 1 + j) - ml)

        def last(j):
            return min(numintervals - 1, (col0 + 1 + j) - ml + k - 1)

        # columns are ordered by start interval (non-decreasing), so the ranges are contiguous
        own_lo = _first(ncols, lambda j: start(j) >= s_lo)
        own_hi = _first(ncols, lambda j: start(j) >= s_hi)
        touch_lo = _first(ncols, lambda j: last(j) >= s_lo)
        touch_hi = own_hi
        out.append(IntervalShard(s_lo, s_hi, own_lo, own_hi, min(touch_lo, own_lo), touch_hi))
    return out


def _first(n, pred):
    """Smallest j in [0, n] with pred(j) true for all j' >= j (monotone predicate)."""
    lo, hi = 0, n
    while lo < hi:
        mid = (lo + hi) // 2
        if pred(mid):
            hi = mid
        else:
            lo = mid + 1
    return lo


def halo_exchange_add(partial: torch.Tensor, shard: IntervalShard, shards: list[IntervalShard], rank: int,
                      group=None) -> torch.Tensor:
    """``partial``: [touch_hi - touch_lo, W] partial sums of this rank (BandedMatrix columns are
    contiguous rows of the (n, W) tensor).  Sends the columns owned by the left neighbour(s) and
    adds what the right neighbour computed for this rank's last columns.  Returns the owned
    columns [own_hi - own_lo, W], complete."""
    world = len(shards)
    W = partial.shape[1]
    own = partial[shard.own_lo - shard.touch_lo:shard.own_hi - shard.touch_lo].clone()
    ops = []
    send_bufs, recv_bufs = [], []
    # columns I touched but that start in an earlier rank's range: their owners are to the left
    for r2 in range(rank):
        lo, hi = max(shards[r2].own_lo, shard.touch_lo), min(shards[r2].own_hi, shard.own_lo)
        if hi > lo:
            buf = partial[lo - shard.touch_lo:hi - shard.touch_lo].contiguous()
            send_bufs.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, _global_rank(r2, group), group=group))
    # columns I own that ranks to the right touched
    for r2 in range(rank + 1, world):
        lo, hi = max(shard.own_lo, shards[r2].touch_lo), min(shard.own_hi, shards[r2].own_lo)
        if hi > lo:
            buf = torch.empty((hi - lo, W), dtype=partial.dtype, device=partial.device)
            recv_bufs.append((lo, hi, buf))
            ops.append(dist.P2POp(dist.irecv, buf, _global_rank(r2, group), group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for lo, hi, buf in recv_bufs:
        own[lo - shard.own_lo:hi - shard.own_lo] += buf
    return own


def _global_rank(r, group):
    return r if group is None else dist.get_global_rank(group, r)


class Comm:
    """``cb_comm`` of this rank: the peer-mapped mailboxes of the fused halo exchange (``csrc/comm.cu``).  With one
    process per GPU the 64-byte handles are all-gathered through ``torch.distributed`` (any backend; the data path
    itself is plain stores over NVLink from inside the assembly kernel)."""

    def __init__(self, rank: int, world: int, group=None, connect: bool = True):
        from . import _lib
        self.rank, self.world = rank, world
        h = C.c_void_p()
        _lib.call("cb_comm_create", rank, world, C.byref(h))
        self.handle = h
        if connect and world > 1:
            buf = (C.c_char * 64)()
            _lib.call("cb_comm_export", self.handle, C.cast(buf, C.c_void_p))
            mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            backend = dist.get_backend(group)
            dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
            parts = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
            dist.all_gather(parts, mine.to(dev), group=group)
            allh = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
            self._handles = C.create_string_buffer(allh, 64 * world)
            _lib.call("cb_comm_connect", self.handle, C.cast(self._handles, C.c_void_p))
            dist.barrier(group=group)               # every mailbox is mapped before anyone pushes

    @staticmethod
    def connect_local(comms):
        """Single-process form: ``comms[r]`` was created (``connect=False``) with the device of rank r current."""
        from . import _lib
        arr = (C.c_void_p * len(comms))(*[c.handle for c in comms])
        _lib.call("cb_comm_connect_local", arr, len(comms))

    def status(self) -> int:
        from . import _lib
        from .operators import _stream
        st = C.c_int(0)
        _lib.call("cb_comm_status", self.handle, C.byref(st), _stream())
        return st.value

    def shard(self, R, rank=None) -> IntervalShard:
        """``cb_interval_shard`` of this (or another) rank for the columns of the (restricted) basis R."""
        from . import _lib
        v = [C.c_int64() for _ in range(5)]
        _lib.call("cb_interval_shard", R.parent_basis.handle, R.col0, R.ncols, self.rank if rank is None else rank, self.world,
                  *[C.byref(x) for x in v])
        s_lo, s_hi, own_lo, own_hi, touch_lo = (x.value for x in v)
        return IntervalShard(s_lo, s_hi, own_lo, own_hi, touch_lo, own_hi)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                from . import _lib
                _lib.load().cb_comm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def assemble_interval_sharded(L, R, a: int, b: int, rank: int, world: int, opvals: torch.Tensor | None = None,
                              coulomb_Z: float | None = None, halo: str = "exchange", group=None, comm: Comm | None = None,
                              out: torch.Tensor | None = None):
    """Interval-sharded ``diffop!`` on this rank's GPU.  Returns (own_lo, own_hi, cols) where ``cols``
    is the [own_hi-own_lo, l+u+1] slice of the BandedMatrix data that this rank owns.

    halo: "fused" — the exchange happens inside the assembly kernel through peer-mapped mailboxes
    (``cb_bspline_assemble_sharded``, needs ``comm``); "recompute" — no communication; "exchange" — the library-external
    baseline: partial bands + one NCCL send/recv per neighbour + add (kept for comparison and for the CPU tests)."""
    from . import _lib
    from .bsplines import band_shape
    from .operators import _ptr, _stream
    PL, PR = L.parent_basis, R.parent_basis
    l, u = band_shape(L, R)
    W = l + u + 1
    nint = PR.t.numintervals()
    shards = interval_shards(nint, PR.k, PR.t.ml, R.col0, R.ncols, world)
    sh = shards[rank]
    dev = PR.device
    cz = 0.0 if coulomb_Z is None else float(coulomb_Z)
    cflag = 0 if coulomb_Z is None else 1
    if comm is not None and halo in ("fused", "recompute"):
        cols = out if out is not None else torch.empty((sh.own_hi - sh.own_lo, W), dtype=torch.float64, device=dev)
        _lib.call("cb_bspline_assemble_sharded", comm.handle, PL.handle, PR.handle, a, b, _ptr(opvals), cflag, cz, L.col0, L.ncols,
                  R.col0, R.ncols, l, u, 0 if halo == "fused" else 1, _ptr(cols), _stream())
        return sh.own_lo, sh.own_hi, cols
    if halo == "fused":
        raise ValueError("halo='fused' needs a Comm")
    if halo == "recompute" or world == 1:
        cols = torch.empty((sh.own_hi - sh.own_lo, W), dtype=torch.float64, device=dev)
        _lib.call("cb_bspline_assemble_part", PL.handle, PR.handle, a, b, _ptr(opvals), cflag, cz, L.col0, L.ncols,
                  R.col0, R.ncols, l, u, sh.own_lo, sh.own_hi, 0, nint, _ptr(cols), _stream())
        return sh.own_lo, sh.own_hi, cols
    if halo != "exchange":
        raise ValueError("halo must be 'exchange' or 'recompute'")
    partial = torch.empty((sh.touch_hi - sh.touch_lo, W), dtype=torch.float64, device=dev)
    _lib.call("cb_bspline_assemble_part", PL.handle, PR.handle, a, b, _ptr(opvals), cflag, cz, L.col0, L.ncols,
              R.col0, R.ncols, l, u, sh.touch_lo, sh.touch_hi, sh.s_lo, sh.s_hi, _ptr(partial), _stream())
    return sh.own_lo, sh.own_hi, halo_exchange_add(partial, sh, shards, rank, group)


def allgather_columns(cols: torch.Tensor, shards: list[IntervalShard], group=None) -> torch.Tensor:
    """Replicates the assembled band on every rank (optional step after a sharded assembly)."""
    world = len(shards)
    W = cols.shape[1]
    parts = [torch.empty((s.own_hi - s.own_lo, W), dtype=cols.dtype, device=cols.device) for s in shards]
    dist.all_gather(parts, cols.contiguous(), group=group) if _equal_sizes(shards) else _allgather_ragged(parts, cols, group)
    return torch.cat(parts, dim=0)


def _equal_sizes(shards):
    return len({s.own_hi - s.own_lo for s in shards}) == 1


def _allgather_ragged(parts, cols, group):
    world = len(parts)
    rank = dist.get_rank(group)
    for r in range(world):
        if r == rank:
            parts[r].copy_(cols)
        dist.broadcast(parts[r], _global_rank(r, group), group=group)


# ---- K4 / K5: FE-DVR element shards and finite-difference node shards ---------------------------------
def fedvr_assemble_sharded(F, nder: int, rank: int, world: int, out: torch.Tensor | None = None):
    """``derop!(A, B, nder)`` for the elements ``batch_range(nel, rank, world)`` of the FE-DVR basis F on this rank's GPU.
    Element blocks only need the (replicated) grid, weights and bridge normalisation of the basis, so there is nothing to
    exchange: the shared bridge entries are sums of two neighbouring corners and are formed when the operator is applied
    or exported.  Returns ``(e_lo, e_hi, blocks)`` with ``blocks`` the slice ``boff[e_lo] : boff[e_hi]`` of the full
    block array."""
    from . import _lib
    from .operators import _ptr, _stream
    P = F.parent_basis
    e_lo, e_hi = batch_range(P.nel, rank, world)
    nloc = e_hi - e_lo
    btot = int(P.boff[-1]) + int(P.order[-1]) ** 2          # boff[e] = sum of order^2 of the elements before e
    b0 = int(P.boff[e_lo]) if e_lo < P.nel else btot
    b1 = int(P.boff[e_hi]) if e_hi < P.nel else btot
    blocks = out if out is not None else torch.empty(b1 - b0, dtype=torch.float64, device=P.device)
    if nloc == 0:
        return e_lo, e_hi, blocks
    if P.uniform_order > 0:
        o = P.uniform_order
        _lib.call("cb_fedvr_assemble", P.d_x.data_ptr() + 8 * e_lo * (o - 1), P.d_wcat.data_ptr() + 8 * e_lo * o,
                  P.d_n.data_ptr() + 8 * e_lo * (o - 1), None, None, None, None, nloc, o, o, nder, _ptr(blocks), _stream())
    else:
        # per-element arrays hold absolute offsets: shift the array pointers by e_lo elements and the block base so that
        # blocks[boff[e_lo]] is the first entry written
        _lib.call("cb_fedvr_assemble", _ptr(P.d_x), _ptr(P.d_wcat), _ptr(P.d_n), P.d_estart.data_ptr() + 8 * e_lo,
                  P.d_order.data_ptr() + 4 * e_lo, P.d_boff.data_ptr() + 8 * e_lo, P.d_woff.data_ptr() + 8 * e_lo, nloc, 0,
                  int(P.order[e_lo:e_hi].max()), nder, blocks.data_ptr() - 8 * b0, _stream())
    return e_lo, e_hi, blocks


def fd_derivative_sharded(R, nder: int, rank: int, world: int, Z: float = 0.0, ell: int = 0):
    """Rows ``batch_range(N, rank, world)`` of the (Sym)Tridiagonal derivative operator of the finite-difference grid R
    (``cb_fd_derivative`` on a node range; stencils depend on the replicated grid only).  Returns
    ``(lo, hi, dl, d, du)``: ``d[i]`` = A[lo+i, lo+i], ``du[i]`` = A[lo+i, lo+i+1], ``dl[i]`` = A[lo+i+1, lo+i]; the
    off-diagonals have ``hi - lo`` entries (one fewer on the last rank)."""
    from . import _lib
    from .operators import _ptr, _stream
    P = R.parent_basis
    N = P.N
    lo, hi = batch_range(N, rank, world)
    n = hi - lo
    ext = n + (1 if hi < N else 0)                  # one more node: the coupling of row hi-1 to row hi
    dev = P.device
    d = torch.empty(max(ext, 1), dtype=torch.float64, device=dev)
    dl = torch.empty(max(ext - 1, 1), dtype=torch.float64, device=dev)
    du = torch.empty(max(ext - 1, 1), dtype=torch.float64, device=dev)
    if n > 0:
        _lib.call("cb_fd_derivative", P.scheme, nder, _ptr(P.d_r), None, N, float(P.step), lo, ext, float(Z), int(ell),
                  _ptr(dl), _ptr(d), _ptr(du), _stream())
    noff = ext - 1 if n > 0 else 0
    return lo, hi, dl[:noff], d[:n], du[:noff]
