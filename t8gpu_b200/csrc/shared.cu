// Cross-process sharing of device buffers between the GPUs of one node (one process per GPU).
// Replaces t8gpu/memory/shared_device_vector.inl:171-198 (cudaIpcGetMemHandle + MPI_Allgather + cudaIpcOpenMemHandle):
// the handle exchange is left to the caller; opened pointers are peer mappings served over NVLink.
#include <cstring>

#include "../../include/t8gpu_b200.h"
#include "common.cuh"
#include "peer_sync.cuh"

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
// Stage barrier + max-reduction between the GPUs of one node over peer memory (one process per GPU); mailbox layout
// and protocol in peer_sync.cuh.  One warp, one lane per rank: every rank stores (value, epoch) into its slot of EVERY
// rank's mailbox through the peer-mapped pointers (NVLink), then waits until all slots of its own mailbox carry the
// epoch.  Kernels of different ranks run on different GPUs, so none of them can keep another from being scheduled.
// Two slot classes with their own epoch sequences: without a value the stage slots (the same ones the stage kernels
// signal through when they order themselves, t8b200_fused_stage_sync_*), with a value the CFL slots.
namespace {
using t8b200::PeerSlot;

__global__ void peer_barrier_kernel(int nranks, int rank, long long epoch, PeerSlot* const* mailboxes, const void* value,
                                    int value_is_f64, void* out_max) {
  t8b200::peer_barrier_warp(threadIdx.x, nranks, rank, epoch, mailboxes, value, value_is_f64, out_max);
}

// CompressibleEulerSolver::compute_timestep (examples/compressible_euler/solver.cu:225-228) without leaving the device:
// dt = cfl * length / vmax, optionally capped; the stage kernels of the next step read it from dt_out.
template <typename T>
__global__ void timestep_kernel(const T* vmax, T cfl, T length, T dt_cap, T* dt_out) {
  T dt = cfl * length / *vmax;
  if (dt_cap > T(0) && !(dt < dt_cap)) dt = dt_cap;   // also catches vmax == 0 (inf) and NaN
  *dt_out = dt;
}
}  // namespace

extern "C" int t8b200_peer_barrier(int nranks, int rank, long long epoch, void* const* mailboxes_dev,
                                   const void* value_dev, int value_is_f64, void* out_max_dev, void* stream) {
  if (nranks < 1 || nranks > 32 || rank < 0 || rank >= nranks || epoch <= 0 || !mailboxes_dev) return cudaErrorInvalidValue;
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(nranks, rank, epoch, (PeerSlot* const*)mailboxes_dev, value_dev,
                                                          value_is_f64, out_max_dev);
  return cudaGetLastError();
}

extern "C" int t8b200_timestep_f32(const float* speed_max_dev, float cfl, float length, float dt_cap, float* dt_dev,
                                   void* stream) {
  if (!speed_max_dev || !dt_dev) return cudaErrorInvalidValue;
  timestep_kernel<float><<<1, 1, 0, (cudaStream_t)stream>>>(speed_max_dev, cfl, length, dt_cap, dt_dev);
  return cudaGetLastError();
}
extern "C" int t8b200_timestep_f64(const double* speed_max_dev, double cfl, double length, double dt_cap, double* dt_dev,
                                   void* stream) {
  if (!speed_max_dev || !dt_dev) return cudaErrorInvalidValue;
  timestep_kernel<double><<<1, 1, 0, (cudaStream_t)stream>>>(speed_max_dev, cfl, length, dt_cap, dt_dev);
  return cudaGetLastError();
}
