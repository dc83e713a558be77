// Small shared helpers for the t8gpu_b200 CUDA translation units.
#pragma once
#include <cuda_runtime.h>

#define T8B_TRY(expr)                      \
  do {                                     \
    cudaError_t _e = (expr);               \
    if (_e != cudaSuccess) return (int)_e; \
  } while (0)
