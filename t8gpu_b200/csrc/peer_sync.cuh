// Stage ordering between the GPUs of one node over peer memory (one process per GPU), shared by the barrier kernel
// (shared.cu) and by the stage kernels (structured.cu, fused.cu), which signal and wait themselves.
//
// Replaces cudaDeviceSynchronize() + MPI_Barrier between the phases of iterate()
// (examples/compressible_euler/solver.cu:98-99,111-112,130-131,143-144,162-163) and the MPI_Allreduce(MAX) of
// compute_timestep (solver.cu:219-223).
//
// Mailbox of rank r: 4 * nranks slots of 16 bytes in r's memory, peer-mapped on every rank.
//   [0, 2n)   stage epochs:  slot (epoch & 1) * n + writer   written by the stage kernels / the barrier kernel
//   [2n, 4n)  CFL values:    slot 2n + (epoch & 1) * n + writer   (value, epoch) written by the barrier kernel
// Slots are double-buffered by epoch parity: a writer can be at most one epoch ahead of a reader (it needs the
// reader's own signal of epoch e to pass e + 1), so the (value, epoch) pair a reader waits for is never overwritten
// before it has been read (one slot per writer allowed that).
#pragma once
#include <cuda_runtime.h>

namespace t8b200 {

struct PeerSlot { double value; long long epoch; };

// What a stage kernel needs to order itself against the peers.  mailboxes == nullptr: off (single rank, or the caller
// orders the stages with t8b200_peer_barrier / NCCL).
struct StageSync {
  PeerSlot* const* mailboxes = nullptr;   // device table, one pointer per rank
  unsigned*        counter   = nullptr;   // partition-boundary chunks of this stage that have finished (this rank)
  long long        wait_epoch = 0;        // > 0: boundary chunks wait until every peer has signalled this epoch
  long long        signal_epoch = 0;      // > 0: the last boundary chunk to finish signals this epoch to every peer
  int              nranks = 0, rank = 0;
  int              n_boundary_total = 0;  // boundary chunks of the whole stage (structured + generic launches)
};

__device__ __forceinline__ void peer_store_epoch(PeerSlot* s, long long epoch) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(&s->epoch), "l"(epoch) : "memory");
}
__device__ __forceinline__ long long peer_load_epoch(const PeerSlot* s) {
  long long e;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(e) : "l"(&s->epoch) : "memory");
  return e;
}

// The barrier protocol of t8b200_peer_barrier for the first `nranks` lanes of ONE warp (all 32 lanes call it): lane p
// stores (value, epoch) into this rank's slot of rank p's mailbox, then waits for rank p's slot of the own mailbox;
// with out_max the maximum of the values goes to *out_max.  value == nullptr: stage slots, else CFL slots.
__device__ __forceinline__ void peer_barrier_warp(int lane, int nranks, int rank, long long epoch, PeerSlot* const* mailboxes,
                                                  const void* value, int value_is_f64, void* out_max) {
  const int base = (value ? 2 * nranks : 0) + (int)(epoch & 1) * nranks;
  double    v    = 0.0;
  if (value) v = value_is_f64 ? *(const double*)value : (double)*(const float*)value;
  if (lane < nranks) {
    PeerSlot* s = mailboxes[lane] + base + rank;
    // value first, then the epoch with release semantics at system scope (the reader acquires the epoch)
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(&s->value), "d"(v) : "memory");
    peer_store_epoch(s, epoch);
  }
  double m = 0.0;
  if (lane < nranks) {
    const PeerSlot* s = mailboxes[rank] + base + lane;
    while (peer_load_epoch(s) < epoch) {}
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(m) : "l"(&s->value) : "memory");
  }
  if (out_max) {
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
      if (value_is_f64) *(double*)out_max = m; else *(float*)out_max = (float)m;
    }
  }
}

// Called by a THREAD about to read an element owned by rank `owner` (a ghost): waits until that rank has signalled
// the awaited epoch.  The acquire is per thread and in front of the thread's own peer loads, so no CTA-wide step is
// needed, threads that read this rank's own elements never wait, and a chunk only waits for the ranks it reads from.
__device__ __forceinline__ void stage_wait_owner(const StageSync& S, int owner) {
  if (S.wait_epoch > 0) {
    const PeerSlot* s = S.mailboxes[S.rank] + (S.wait_epoch & 1) * S.nranks + owner;
    while (peer_load_epoch(s) < S.wait_epoch) {}
  }
}

// Called by ONE thread of a boundary chunk after a CTA barrier that follows the chunk's last global store: counts the
// chunk; the last one publishes the epoch to every peer.  Ordering: every chunk releases its stores at GPU scope
// before it counts (fence + atomic), the last chunk acquires the count and releases at SYSTEM scope in front of the
// flag -- causality order is transitive, so a peer that acquires the flag sees the stores of all chunks, while only
// one thread per stage pays for a system-scope fence.
__device__ __forceinline__ void stage_signal(const StageSync& S) {
  __threadfence();
  const unsigned done = atomicAdd(S.counter, 1u);
  if (done + 1u == (unsigned)S.n_boundary_total) {
    *S.counter = 0u;   // every boundary chunk of this stage has counted; the next stage starts from zero
    __threadfence_system();
    for (int r = 0; r < S.nranks; r++)
      if (r != S.rank) peer_store_epoch(S.mailboxes[r] + (S.signal_epoch & 1) * S.nranks + S.rank, S.signal_epoch);
  }
}

}  // namespace t8b200
