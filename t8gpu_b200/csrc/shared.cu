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
