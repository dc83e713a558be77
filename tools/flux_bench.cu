// Micro-benchmark: the face flux in isolation (cells from shared memory, fluxes to shared memory, as in phase 1 of
// fused_stage_kernel) at different register caps and warps per SM -> cycles per face per warp and FP64 pipe use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I t8gpu_b200/csrc -o tools/flux_bench tools/flux_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "euler_flux.cuh"
using namespace t8b200;

template <int MINB, int ILP, bool REC = false>
__global__ void __launch_bounds__(256, MINB) flux_loop(double* out, int iters, long long* cyc, const unsigned* rec = nullptr) {
  extern __shared__ double sm[];
  double* cq = sm;               // [7][512]
  double* fl = sm + 7 * 512;     // [5][1024]
  const int tid = threadIdx.x;
  for (int s = tid; s < 512; s += 256) {
    Cell<double> c = to_cell(1.0 + 1e-3 * (s % 17), 0.1 + 1e-3 * (s % 5), 0.05, -0.02, 2.5 + 1e-3 * (s % 7));
    cq[s] = c.rho; cq[512 + s] = c.hx; cq[1024 + s] = c.hy; cq[1536 + s] = c.hz; cq[2048 + s] = c.kp; cq[2560 + s] = c.b; cq[3072 + s] = c.q;
  }
  __syncthreads();
  auto ld = [&](int s) { Cell<double> c; c.rho = cq[s]; c.hx = cq[512 + s]; c.hy = cq[1024 + s]; c.hz = cq[1536 + s]; c.kp = cq[2048 + s]; c.b = cq[2560 + s]; c.q = cq[3072 + s]; return c; };
  double smax = 0;
  unsigned lr_n = REC ? rec[(blockIdx.x & 1023) * 1024 + tid] : 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < ILP; u++) {
      int l = (tid + it * 3 + u * 64) & 511, r = (l + 1 + (it & 7)) & 511;
      if (REC) {   // as the kernel: the record of the next face is requested while this one is evaluated
        const unsigned lr = lr_n;
        lr_n = rec[((blockIdx.x + it) & 1023) * 1024 + ((tid + 256 * (it + 1)) & 1023)];
        l = lr & 511; r = (lr >> 16) & 511;
      }
      const Cell<double> L = ld(l), R = ld(r);
      double F[5];
      const double s = kepes_flux_n<double, 0>(L, R, 0.0, 0.0, 0.0, F);
      smax = fmax_(smax, s);
      const int j = (tid + u * 256 + (it & 1) * 512) & 1023;
#pragma unroll
      for (int k = 0; k < 5; k++) fl[k * 1024 + j] = F[k];
    }
  }
  const long long t1 = clock64();
  if (tid == 0) atomicAdd((unsigned long long*)cyc, (unsigned long long)(t1 - t0));
  if (smax == 123.456) out[0] = smax + fl[tid];
}

template <int MINB, int ILP, bool REC = false>
void run(int ctas_per_sm, int sms) {
  double* d; long long* c;
  cudaMalloc(&d, 8); cudaMalloc(&c, 8); cudaMemset(c, 0, 8);
  const size_t smem = 8 * (7 * 512 + 5 * 1024);
  auto k = flux_loop<MINB, ILP, REC>;
  static unsigned* rec = nullptr;
  if (!rec) {   // face records: left slots consecutive, right slots = a Morton-like neighbour (as in a hex chunk)
    unsigned* h = new unsigned[1024 * 1024];
    for (int i = 0; i < 1024 * 1024; i++) {
      const int j = i & 1023, l = j & 255, ax = (j >> 8) % 3;
      const int r = ax == 0 ? (l ^ 1) : ax == 1 ? (l ^ 2) : (l ^ 4);
      h[i] = (unsigned)l | ((unsigned)((r + ((j & 7) == 0 ? 256 : 0)) & 511) << 16);
    }
    cudaMalloc(&rec, 4u << 20); cudaMemcpy(rec, h, 4u << 20, cudaMemcpyHostToDevice); delete[] h;
  }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
  const int iters = 2048 / ILP, blocks = sms * ctas_per_sm;
  k<<<blocks, 256, smem>>>(d, 16, c, rec);
  cudaMemset(c, 0, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<<<blocks, 256, smem>>>(d, iters, c, rec);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
  const double faces = (double)blocks * 256 * iters * ILP;
  printf("%s regs %3d ILP %d CTAs/SM %d (%2d warps/SM): %7.1f cycles per face per warp, %6.2f G faces/s, ~%4.1f%% of FP64 peak (113 FP64/face)\n",
         REC ? "records" : "synthetic", fa.numRegs, ILP, ctas_per_sm, ctas_per_sm * 8, (double)cy / blocks / (iters * ILP), faces / ms * 1e-6,
         100.0 * faces * 113 / (ms * 1e-3) / 16.5e12);
  cudaFree(d); cudaFree(c);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs; error %s\n", p.name, sms, cudaGetErrorString(cudaGetLastError()));
  run<3, 1>(1, sms); run<3, 1>(2, sms); run<3, 1>(3, sms);
  run<3, 1, true>(1, sms); run<3, 1, true>(2, sms); run<3, 1, true>(3, sms);
  printf("error %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
