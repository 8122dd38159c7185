// asm_halo.cuh — mailbox protocol of the fused halo exchange (shared by the assembly kernels; comm.cu fills it).
#pragma once
#include "common.cuh"

// Fused halo exchange of an interval-sharded assembly.  A rank sums its own intervals into every
// column that touches them.  Its first n_push columns belong to the LEFT neighbour: they are stored straight into that
// neighbour's mailbox over NVLink (peer-mapped memory) and, once every tile that holds such a column has finished, a
// flag is raised there.  Its last n_recv columns also live on the RIGHT neighbour's intervals: the tiles that hold them
// wait for the local flag (the right neighbour handles its pushed tiles first, these tiles come last, so the wait is
// normally over before it starts) and add the mailbox to their own sums.  No host round trip, no collective launch.
// Mailboxes are double-buffered by call parity; `ack` words tell the pusher that the buffer of two calls ago has been
// consumed.  Every spin is bounded (status word).
struct AsmHalo {
    int n_push, n_recv;
    int n_push_tiles, n_recv_tiles;
    double *push_cols;                      // peer: left neighbour's mailbox columns of this parity
    unsigned long long *push_flag;          // peer: its flag of this parity
    const double *recv_cols;                // local mailbox columns of this parity
    unsigned long long *recv_flag;          // local flag of this parity (written by the right neighbour)
    unsigned long long *ack_out;            // peer: right neighbour's ack word of this parity (we write the epoch after consuming)
    unsigned long long *ack_in;             // local ack word of this parity (the left neighbour consumed that epoch)
    unsigned long long epoch;
    unsigned int *counters;                 // local: [0] pushed tiles finished, [1] receiving tiles finished
    int *status;                            // local: non-zero after a spin timeout
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// waits until *p >= want; gives up after ~15 s (a peer that never launched must not hang the GPU for good, but ranks of one
// job may legitimately be seconds apart: host-side work between calls, first-call initialisation)
__device__ __forceinline__ void halo_spin(const unsigned long long *p, unsigned long long want, int *status) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < want) {
        if (clock64() - t0 > 30000000000ll) { atomicExch(status, 1); break; }
        __nanosleep(64);
    }
}

