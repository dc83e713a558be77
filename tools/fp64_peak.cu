// Micro-benchmark: FP64 FMA issue rate per SM (to place the FP64-pipe roofline next to the HBM one).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void __launch_bounds__(256) dfma(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  if (s == 12345.678) out[0] = s;
}
template <int ILP>
void run(int warps_per_sm, int sms) {
  double* d;
  cudaMalloc(&d, 8);
  int threads = 256, iters = 20000;
  int blocks = sms * warps_per_sm * 32 / threads;
  if (blocks < 1) { blocks = sms; threads = warps_per_sm * 32; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma<ILP><<<blocks, threads>>>(d, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  dfma<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fmas = (double)blocks * threads * iters * ILP;
  printf("ILP %d warps/SM %2d: %.2f TFMA/s = %.1f TFLOP/s, %.1f FMA lanes/clk/SM @1.965GHz\n", ILP, warps_per_sm,
         fmas / ms * 1e-9, 2 * fmas / ms * 1e-9, fmas / (ms * 1e-3) / sms / 1.965e9);
  cudaFree(d);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  for (int w : {4, 8, 16, 32, 64}) { run<1>(w, sms); run<2>(w, sms); run<4>(w, sms); run<8>(w, sms); }
  return 0;
}
