// Small shared helpers for the t8gpu_b200 CUDA translation units.
#pragma once
#include <cuda_runtime.h>

#include <new>

#define T8B_TRY(expr)                      \
  do {                                     \
    cudaError_t _e = (expr);               \
    if (_e != cudaSuccess) return (int)_e; \
  } while (0)

// scratch device allocation that is released on every exit path
template <typename P>
struct T8bScratch {
  P* p = nullptr;
  T8bScratch() = default;
  T8bScratch(const T8bScratch&)            = delete;
  T8bScratch& operator=(const T8bScratch&) = delete;
  ~T8bScratch() { cudaFree(p); }
};
