// comm.cu — multi-GPU plumbing that a `ccall` host can reach (SURVEY §8b "multi-GPU", §8e; no reference counterpart:
// the reference is single-process and single-threaded).
//
// Only the interval-sharded band assembly has an exchange step (the k-1 boundary columns of neighbouring interval
// ranges); point, vector and pair batches are split with no communication (cb_batch_range).  The exchange is fused
// into the assembly kernel (assemble.cu, AsmHalo): a rank's kernel stores the partial columns it computed for its
// left neighbour straight into that neighbour's mailbox — device memory of the peer GPU mapped into this process,
// reached over NVLink/NVSwitch with ordinary stores — and raises a flag there; the owner's last tiles add the mailbox.
//
// Peer mapping works both ways round:
//   * one process per GPU (torchrun, MPI.jl): cb_comm_create, exchange the 64-byte handles of cb_comm_export with
//     whatever the host has (torch.distributed / MPI all-gather), cb_comm_connect  -> cudaIpcOpenMemHandle;
//   * one process driving several GPUs (a Julia session with CUDA.devices()): cb_comm_create once per device, then
//     cb_comm_connect_local with the array of handles -> cudaDeviceEnablePeerAccess, plain pointers.
#include "common.cuh"

// keep in sync with AsmHalo in assemble.cu (passed through a void* to avoid a shared header for one struct)
struct AsmHaloView {
    int n_push, n_recv;
    int n_push_tiles, n_recv_tiles;
    double *push_cols;
    unsigned long long *push_flag;
    const double *recv_cols;
    unsigned long long *recv_flag;
    unsigned long long *ack_out;
    unsigned long long *ack_in;
    unsigned long long epoch;
    unsigned int *counters;
    int *status;
};

int cb_bspline_assemble_halo_internal(const cb_bspline *L, const cb_bspline *R, int a, int b, const double *opvals,
                                      int coulomb, double Z, i64 rowL0, i64 m, i64 colR0, i64 n, int l, int u,
                                      i64 col_lo, i64 col_hi, i64 s_lo, i64 s_hi, double *band, void *stream,
                                      const void *halo);   // assemble.cu

constexpr int COMM_MAXCOLS = CB_KMAX - 1;                  // boundary columns of a shard
constexpr int COMM_MAXW = 2 * CB_KMAX - 1;                 // band rows
constexpr int COMM_COLS = COMM_MAXCOLS * COMM_MAXW;        // doubles per mailbox buffer

// One rank's mailbox, in ITS device memory; neighbours write into it through their mapping.
struct CommSlot {
    unsigned long long flag[2];                            // [parity] epoch of the columns the right neighbour pushed
    unsigned long long ack[2];                             // [parity] epoch the left neighbour has consumed
    unsigned int counters[2];
    int status;
    int pad;
    double cols[2][COMM_COLS];
};

struct cb_comm_s {
    int rank, world, device;
    CommSlot *slot;                                        // local
    CommSlot *left, *right;                                // neighbours' slots mapped here (NULL at the ends)
    bool left_ipc, right_ipc;
    unsigned long long epoch;                              // sharded calls issued so far (same on every rank)
};

extern "C" {

/* Contiguous share [lo, hi) of n independent items for `rank` of `world` (points, vectors, pairs, intervals). */
int cb_batch_range(cb_i64 n, int rank, int world, cb_i64 *lo, cb_i64 *hi) {
    CB_REQUIRE(n >= 0 && world >= 1 && rank >= 0 && rank < world && lo && hi, "cb_batch_range: bad argument");
    const i64 base = n / world, rem = n % world;
    *lo = rank * base + (rank < rem ? rank : rem);
    *hi = *lo + base + (rank < rem ? 1 : 0);
    return CB_OK;
}

int cb_comm_create(int rank, int world, cb_comm **out) {
    CB_REQUIRE(out != nullptr, "cb_comm_create: out is NULL");
    *out = nullptr;
    CB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "cb_comm_create: rank %d outside 0..%d", rank, world - 1);
    cb_comm_s *c = new cb_comm_s();
    c->rank = rank; c->world = world; c->slot = c->left = c->right = nullptr; c->left_ipc = c->right_ipc = false; c->epoch = 0;
    cudaError_t e = cudaGetDevice(&c->device);
    if (e == cudaSuccess) e = cudaMalloc(&c->slot, sizeof(CommSlot));   // cudaMalloc, not the stream-ordered pool: IPC-exportable
    if (e == cudaSuccess) e = cudaMemset(c->slot, 0, sizeof(CommSlot));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cb_set_error("cb_comm_create: %s", cudaGetErrorString(e)); cudaFree(c->slot); delete c; return CB_ERR_CUDA; }
    *out = c;
    return CB_OK;
}

/* 64-byte handle of this rank's mailbox (cudaIpcMemHandle_t), to be all-gathered by the host. */
int cb_comm_export(const cb_comm *c, void *handle64) {
    CB_REQUIRE(c != nullptr && handle64 != nullptr, "cb_comm_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    CB_CUDA(cudaIpcGetMemHandle(&h, c->slot));
    memcpy(handle64, &h, 64);
    return CB_OK;
}

/* all_handles: world x 64 bytes in rank order (the all-gather of cb_comm_export).  Maps the two neighbours' mailboxes. */
int cb_comm_connect(cb_comm *c, const void *all_handles) {
    CB_REQUIRE(c != nullptr && all_handles != nullptr, "cb_comm_connect: NULL argument");
    const char *hs = static_cast<const char *>(all_handles);
    for (int side = 0; side < 2; ++side) {
        const int nb = side == 0 ? c->rank - 1 : c->rank + 1;
        if (nb < 0 || nb >= c->world) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + 64 * (size_t)nb, 64);
        void *p = nullptr;
        CB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        if (side == 0) { c->left = static_cast<CommSlot *>(p); c->left_ipc = true; }
        else { c->right = static_cast<CommSlot *>(p); c->right_ipc = true; }
    }
    return CB_OK;
}

/* Single-process form: comms[r] was created with the device of rank r current. */
int cb_comm_connect_local(cb_comm **comms, int world) {
    CB_REQUIRE(comms != nullptr && world >= 1, "cb_comm_connect_local: bad argument");
    int cur = 0;
    CB_CUDA(cudaGetDevice(&cur));
    for (int r = 0; r < world; ++r) {
        CB_REQUIRE(comms[r] != nullptr && comms[r]->rank == r && comms[r]->world == world, "cb_comm_connect_local: comms[%d] is not rank %d of %d", r, r, world);
        for (int nb = r - 1; nb <= r + 1; nb += 2) {
            if (nb < 0 || nb >= world) continue;
            if (comms[nb]->device != comms[r]->device) {
                int can = 0;
                CB_CUDA(cudaDeviceCanAccessPeer(&can, comms[r]->device, comms[nb]->device));
                CB_REQUIRE(can, "cb_comm_connect_local: device %d cannot access device %d", comms[r]->device, comms[nb]->device);
                CB_CUDA(cudaSetDevice(comms[r]->device));
                cudaError_t e = cudaDeviceEnablePeerAccess(comms[nb]->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                CB_CUDA(e);
            }
            (nb < r ? comms[r]->left : comms[r]->right) = comms[nb]->slot;
        }
    }
    CB_CUDA(cudaSetDevice(cur));
    return CB_OK;
}

int cb_comm_destroy(cb_comm *c) {
    if (!c) return CB_OK;
    if (c->left && c->left_ipc) cudaIpcCloseMemHandle(c->left);
    if (c->right && c->right_ipc) cudaIpcCloseMemHandle(c->right);
    cudaFree(c->slot);
    delete c;
    return CB_OK;
}

/* Non-zero if a kernel of this rank gave up waiting for a neighbour (its result is then incomplete).  Synchronises
 * `stream` first. */
int cb_comm_status(const cb_comm *c, int *status, void *stream) {
    CB_REQUIRE(c != nullptr && status != nullptr, "cb_comm_status: NULL argument");
    CB_CUDA(cudaMemcpyAsync(status, &c->slot->status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return CB_OK;
}

/* What rank `rank` of `world` computes in an interval-sharded assembly of columns colR0+1 .. colR0+n of R (0-based,
 * half-open; columns are relative to the restriction):
 *   s_lo, s_hi      t.t intervals summed by the rank
 *   own_lo, own_hi  columns whose first interval lies in [s_lo, s_hi): complete on this rank after the exchange
 *   touch_lo        first column with any support in [s_lo, s_hi): columns touch_lo .. own_lo-1 are pushed left
 * Function f (1-based, order k) lives on the intervals f-ml .. f-ml+k-1 clipped to 0 .. numintervals-1
 * (src/knot_sets.jl:138-151). */
int cb_interval_shard(const cb_bspline *R, cb_i64 colR0, cb_i64 n, int rank, int world,
                      cb_i64 *s_lo, cb_i64 *s_hi, cb_i64 *own_lo, cb_i64 *own_hi, cb_i64 *touch_lo) {
    CB_REQUIRE(R != nullptr && s_lo && s_hi && own_lo && own_hi && touch_lo, "cb_interval_shard: NULL argument");
    CB_REQUIRE(colR0 >= 0 && n >= 0 && colR0 + n <= R->nf, "cb_interval_shard: restriction outside 1..nf");
    const i64 nint = R->n_tt - 1;
    int rc = cb_batch_range(nint, rank, world, s_lo, s_hi);
    if (rc) return rc;
    const int k = R->k, ml = R->ml;
    auto start = [&](i64 j) { const i64 s = (colR0 + 1 + j) - ml; return s < 0 ? (i64)0 : s; };
    auto last = [&](i64 j) { const i64 s = (colR0 + 1 + j) - ml + k - 1; return s > nint - 1 ? nint - 1 : s; };
    auto first_true = [&](auto pred) {              // smallest j in [0, n] from which the monotone predicate holds
        i64 lo = 0, hi = n;
        while (lo < hi) { const i64 mid = (lo + hi) / 2; if (pred(mid)) hi = mid; else lo = mid + 1; }
        return lo;
    };
    *own_lo = first_true([&](i64 j) { return start(j) >= *s_lo; });
    *own_hi = first_true([&](i64 j) { return start(j) >= *s_hi; });
    const i64 tl = first_true([&](i64 j) { return last(j) >= *s_lo; });
    *touch_lo = tl < *own_lo ? tl : *own_lo;
    return CB_OK;
}

/* Interval-sharded diffop! / overlap_matrix! on this rank (SURVEY §8e K3 + K3h).  own_cols receives the columns
 * own_lo .. own_hi-1 of the BandedMatrix data ((l+u+1) x (own_hi-own_lo), column-major; cb_interval_shard gives the
 * range).  halo: 0 = exchange — every rank sums its own intervals only and the boundary columns travel through the
 * peer-mapped mailboxes inside the kernel; 1 = recompute — the rank sums its own columns over all their intervals
 * (no communication; (k-1) extra intervals of work).  Every rank of the communicator must make the same sequence of
 * exchange calls.  coulomb: 0 = opvals / none, 1 = op(x) = -Z/x. */
int cb_bspline_assemble_sharded(cb_comm *c, const cb_bspline *L, const cb_bspline *R, int a, int b,
                                const double *opvals, int coulomb, double Z,
                                cb_i64 rowL0, cb_i64 m, cb_i64 colR0, cb_i64 n, int l, int u,
                                int halo, double *own_cols, void *stream) {
    CB_REQUIRE(c != nullptr && R != nullptr && L != nullptr, "cb_bspline_assemble_sharded: NULL argument");
    CB_REQUIRE(halo == 0 || halo == 1, "cb_bspline_assemble_sharded: halo must be 0 (exchange) or 1 (recompute)");
    i64 s_lo, s_hi, own_lo, own_hi, touch_lo;
    int rc = cb_interval_shard(R, colR0, n, c->rank, c->world, &s_lo, &s_hi, &own_lo, &own_hi, &touch_lo);
    if (rc) return rc;
    const i64 nint = R->n_tt - 1;
    const int W = l + u + 1;
    if (halo == 1 || c->world == 1) {
        if (own_hi == own_lo) return CB_OK;
        return cb_bspline_assemble_halo_internal(L, R, a, b, opvals, coulomb, Z, rowL0, m, colR0, n, l, u, own_lo, own_hi, 0, nint,
                                                 own_cols, stream, nullptr);
    }
    // the boundary columns of a rank must come from its direct neighbour only
    CB_REQUIRE(nint / c->world >= R->k, "cb_bspline_assemble_sharded: fewer than k = %d intervals per rank", R->k);
    CB_REQUIRE(W <= COMM_MAXW, "cb_bspline_assemble_sharded: band of %d rows exceeds the mailbox", W);
    i64 r_touch_lo = own_hi;                        // what the right neighbour pushes to us
    if (c->rank + 1 < c->world) {
        i64 t0, t1, t2, t3;
        rc = cb_interval_shard(R, colR0, n, c->rank + 1, c->world, &t0, &t1, &t2, &t3, &r_touch_lo);
        if (rc) return rc;
    }
    AsmHaloView H;
    memset(&H, 0, sizeof(H));
    H.n_push = (int)(own_lo - touch_lo);
    H.n_recv = (int)(own_hi - r_touch_lo);
    CB_REQUIRE(H.n_push >= 0 && H.n_push <= COMM_MAXCOLS && H.n_recv >= 0 && H.n_recv <= COMM_MAXCOLS && H.n_recv <= own_hi - own_lo,
               "cb_bspline_assemble_sharded: boundary of %d / %d columns does not fit", H.n_push, H.n_recv);
    CB_REQUIRE((H.n_push == 0 || c->left != nullptr) && (H.n_recv == 0 || c->right != nullptr),
               "cb_bspline_assemble_sharded: the communicator is not connected (cb_comm_connect)");
    const unsigned long long epoch = ++c->epoch;
    const int par = (int)(epoch & 1);
    H.epoch = epoch;
    if (H.n_push > 0) { H.push_cols = c->left->cols[par]; H.push_flag = &c->left->flag[par]; H.ack_in = &c->slot->ack[par]; }
    if (H.n_recv > 0) { H.recv_cols = c->slot->cols[par]; H.recv_flag = &c->slot->flag[par]; H.ack_out = &c->right->ack[par]; }
    H.counters = c->slot->counters;
    H.status = &c->slot->status;
    if (own_hi == touch_lo) return CB_OK;
    // band base such that column own_lo lands at own_cols[0] (the pushed columns before it never touch this pointer)
    double *base = own_cols - (i64)W * H.n_push;
    return cb_bspline_assemble_halo_internal(L, R, a, b, opvals, coulomb, Z, rowL0, m, colR0, n, l, u, touch_lo, own_hi, s_lo, s_hi,
                                             base, stream, &H);
}

}  // extern "C"
