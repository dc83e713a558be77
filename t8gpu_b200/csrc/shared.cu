// Cross-process sharing of device buffers between the GPUs of one node (one process per GPU).
// Replaces t8gpu/memory/shared_device_vector.inl:171-198 (cudaIpcGetMemHandle + MPI_Allgather + cudaIpcOpenMemHandle):
// the handle exchange is left to the caller; opened pointers are peer mappings served over NVLink.
#include <cstring>

#include "../../include/t8gpu_b200.h"
#include "common.cuh"

extern "C" {

int t8b200_shared_alloc(size_t bytes, void** dev_ptr, unsigned char handle[64]) {
  if (!dev_ptr || !handle || bytes == 0) return cudaErrorInvalidValue;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  void* p = nullptr;
  T8B_TRY(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);   // the reference leaves new allocations uninitialised (SURVEY D-14)
  if (e != cudaSuccess) { cudaFree(p); return e; }
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return e; }
  memcpy(handle, &h, 64);
  *dev_ptr = p;
  return 0;
}

int t8b200_shared_open(const unsigned char handle[64], void** dev_ptr) {
  if (!dev_ptr || !handle) return cudaErrorInvalidValue;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  return cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int t8b200_shared_close(void* dev_ptr) { return dev_ptr ? cudaIpcCloseMemHandle(dev_ptr) : 0; }
int t8b200_shared_free(void* dev_ptr) { return dev_ptr ? cudaFree(dev_ptr) : 0; }
}

// ---------------------------------------------------------------------------------------------------------------
// Stage barrier + max-reduction between the GPUs of one node over peer memory (one process per GPU).
// Replaces cudaDeviceSynchronize() + MPI_Barrier between the phases of iterate() (examples/compressible_euler/
// solver.cu:98-99, ...) and the MPI_Allreduce(MAX) of compute_timestep (solver.cu:219-223) without leaving the stream:
// every rank stores (value, epoch) into its slot of EVERY rank's mailbox through the peer-mapped pointers (NVLink),
// then waits until all slots of its own mailbox carry the epoch.  One warp, one lane per rank.  Kernels of different
// ranks run on different GPUs, so none of them can keep another from being scheduled.
namespace {
struct Slot { double value; long long epoch; };

__global__ void peer_barrier_kernel(int nranks, int rank, long long epoch, Slot* const* mailboxes, const void* value,
                                    int value_is_f64, void* out_max) {
  const int lane = threadIdx.x;
  double    v    = 0.0;
  if (value) v = value_is_f64 ? *(const double*)value : (double)*(const float*)value;
  if (lane < nranks) {
    Slot* s = mailboxes[lane] + rank;
    // value first, then the epoch with release semantics at system scope (the reader acquires the epoch)
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(&s->value), "d"(v) : "memory");
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(&s->epoch), "l"(epoch) : "memory");
  }
  double m = 0.0;
  if (lane < nranks) {
    const Slot* s = mailboxes[rank] + lane;
    long long   e;
    do {
      asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(e) : "l"(&s->epoch) : "memory");
    } while (e < epoch);
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(m) : "l"(&s->value) : "memory");
  }
  if (out_max) {
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
      if (value_is_f64) *(double*)out_max = m; else *(float*)out_max = (float)m;
    }
  }
}
}  // namespace

extern "C" int t8b200_peer_barrier(int nranks, int rank, long long epoch, void* const* mailboxes_dev,
                                   const void* value_dev, int value_is_f64, void* out_max_dev, void* stream) {
  if (nranks < 1 || nranks > 32 || rank < 0 || rank >= nranks || epoch <= 0 || !mailboxes_dev) return cudaErrorInvalidValue;
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(nranks, rank, epoch, (Slot* const*)mailboxes_dev, value_dev,
                                                          value_is_f64, out_max_dev);
  return cudaGetLastError();
}
