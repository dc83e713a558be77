/// @file cuda.h  --  t8gpu_b200 mirror of t8gpu/utils/cuda.h:7-33 (same macro names and abort-on-error behaviour).
#ifndef T8GPU_B200_UTILS_CUDA_H
#define T8GPU_B200_UTILS_CUDA_H

#include <cuda_runtime.h>
#include <sc.h>

#include <cstdio>

namespace t8gpu::detail {
  inline void cuda_fail(cudaError_t e, char const* file, int line) {
    std::fprintf(stderr, "t8gpu: CUDA error %d (%s) at %s:%d\n", static_cast<int>(e), cudaGetErrorString(e), file, line);
    SC_ABORT("CUDA error caught");
  }
}  // namespace t8gpu::detail

/// Evaluates a CUDA runtime call (or a t8b200_* C-ABI call, which returns a cudaError_t as int) and aborts on error.
#define T8GPU_CUDA_CHECK_ERROR(expr)                                                                   \
  do {                                                                                                 \
    cudaError_t t8gpu_err_ = static_cast<cudaError_t>(expr);                                           \
    if (t8gpu_err_ != cudaSuccess) ::t8gpu::detail::cuda_fail(t8gpu_err_, __FILE__, __LINE__);         \
  } while (0)

/// After a kernel launch: always checks the launch status; debug builds also synchronise (reference: cuda.h:20-33).
#ifdef NDEBUG
#define T8GPU_CUDA_CHECK_LAST_ERROR() T8GPU_CUDA_CHECK_ERROR(cudaGetLastError())
#else
#define T8GPU_CUDA_CHECK_LAST_ERROR()              \
  do {                                             \
    T8GPU_CUDA_CHECK_ERROR(cudaGetLastError());    \
    T8GPU_CUDA_CHECK_ERROR(cudaDeviceSynchronize()); \
  } while (0)
#endif

#endif  // T8GPU_B200_UTILS_CUDA_H
