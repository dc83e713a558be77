/// @file profiling.h  --  same macro names as t8gpu/utils/profiling.h:7-36; implemented with CUDA events on the
/// default stream (device time) instead of a host wall clock, because the B200 path is asynchronous.
#ifndef T8GPU_B200_UTILS_PROFILING_H
#define T8GPU_B200_UTILS_PROFILING_H

#include <cuda_runtime.h>

#include <cstdio>

namespace t8gpu::detail {
  struct EventTimer {
    cudaEvent_t a{}, b{};
    char const* name;
    explicit EventTimer(char const* n) : name{n} {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, 0);
    }
    void stop() {
      cudaEventRecord(b, 0);
      cudaEventSynchronize(b);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, a, b);
      std::fprintf(stderr, "[t8gpu] %s: %.3f ms\n", name, ms);
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  };
}  // namespace t8gpu::detail

#define T8GPU_TIMER_START(name) ::t8gpu::detail::EventTimer t8gpu_timer_##name{#name}
#define T8GPU_TIMER_STOP(name) t8gpu_timer_##name.stop()
#define T8GPU_TIME(expr)                            \
  do {                                              \
    ::t8gpu::detail::EventTimer t8gpu_timer_{#expr}; \
    expr;                                           \
    t8gpu_timer_.stop();                            \
  } while (0)

#endif  // T8GPU_B200_UTILS_PROFILING_H
