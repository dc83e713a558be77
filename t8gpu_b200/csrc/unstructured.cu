// Unstructured (MeshManager) hot path for sm_100a, reference-shaped entry points: face flux, SSP-RK3 stage,
// wave-speed reduction (the fused tile-plan stage kernel is in fused.cu).  C ABI in include/t8gpu_b200.h.
//
// Reference behaviour replaced (not translated):
//   examples/compressible_euler/kernels.cu:135-469, solver.cu:75-229, t8gpu/timestepping/ssp_runge_kutta.inl:30-99
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/t8gpu_b200.h"
#include "common.cuh"
#include "euler_flux.cuh"

using namespace t8b200;

// ============================================================================================================
// 1. reference-shaped kernels
// ============================================================================================================

template <typename T>
struct TablesC { const T* const* p[5]; };
template <typename T>
struct Tables { T* const* p[5]; };
template <typename T, int N = 5>
struct PtrsC { const T* p[N]; };
template <typename T, int N = 5>
struct Ptrs { T* p[N]; };

// One thread per face (interior faces first, then boundary faces).  The gathers are indirect by nature; what we
// fix relative to the reference is the arithmetic (euler_flux.cuh) and the 10 pointer-table reloads per side
// (tables are read once into registers: they are warp-uniform and L1-resident).
template <typename T>
__global__ void __launch_bounds__(256)
flux_faces_kernel(int nf, int nb, const int32_t* __restrict__ ranks, const int32_t* __restrict__ indices,
                  const int32_t* __restrict__ nbr, const T* __restrict__ normals, const T* __restrict__ areas,
                  TablesC<T> u, Tables<T> fl, T* __restrict__ speed) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf + nb) return;
  const bool boundary = f >= nf;
  int        l = boundary ? nbr[2 * nf + (f - nf)] : nbr[2 * f];
  int        r = boundary ? l : nbr[2 * f + 1];
  int lr = 0, li = l, rr = 0, ri = r;
  if (ranks) {
    lr = ranks[l];
    li = indices[l];
    rr = ranks[r];
    ri = indices[r];
  }
  T nx = normals[3 * f], ny = normals[3 * f + 1], nz = normals[3 * f + 2];
  T a  = areas[f];

  Cell<T> L = to_cell(u.p[0][lr][li], u.p[1][lr][li], u.p[2][lr][li], u.p[3][lr][li], u.p[4][lr][li]);
  Cell<T> R = boundary ? mirror(L, nx, ny, nz)
                       : to_cell(u.p[0][rr][ri], u.p[1][rr][ri], u.p[2][rr][ri], u.p[3][rr][ri], u.p[4][rr][ri]);
  T F[5];
  T s = kepes_flux(L, R, nx, ny, nz, F);
  if (speed) speed[f] = s;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    T v = a * F[k];
    atomicAdd(&fl.p[k][lr][li], -v);
    if (!boundary) atomicAdd(&fl.p[k][rr][ri], v);
  }
}

template <typename T, int STAGE>
__global__ void __launch_bounds__(256)
rk3_stage_kernel(int64_t n, int nvar, PtrsC<T, 8> prev, PtrsC<T, 8> in, Ptrs<T, 8> out, Ptrs<T, 8> flux,
                 const T* __restrict__ vol, int cells_per_vol, T dt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T v = vol[i / cells_per_vol];
  if (cells_per_vol > 1) v = v / T(cells_per_vol);
#pragma unroll
  for (int k = 0; k < 8; k++) {  // unrolled with a predicate so the pointer arrays stay in registers / constant bank
    if (k < nvar) {
      T p = prev.p[k][i];
      T s = STAGE == 1 ? T(0) : in.p[k][i];
      out.p[k][i]  = rk_combine<T, STAGE>(p, s, flux.p[k][i], dt, v);
      flux.p[k][i] = T(0);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) max_speed_kernel(const T* __restrict__ speed, int64_t n, T* out) {
  T       m      = T(0);
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = fmax_(m, speed[i]);
  m = warp_max(m);
  __shared__ T sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = sm[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) {
      T w = __shfl_xor_sync(0xffu, m, o);
      m   = m > w ? m : w;
    }
    if (threadIdx.x == 0) atomic_max_nonneg(out, m);
  }
}

template <typename T>
static int flux_faces_impl(int32_t nf, int32_t nb, const int32_t* ranks, const int32_t* indices, const int32_t* nbr,
                           const T* normals, const T* areas, const T* const* const* vars_all,
                           T* const* const* flux_all, T* speed, void* stream) {
  if (nf < 0 || nb < 0 || !vars_all || !flux_all) return cudaErrorInvalidValue;
  if (nf + nb == 0) return cudaSuccess;
  if (!nbr || !normals || !areas) return cudaErrorInvalidValue;
  TablesC<T> u;
  Tables<T>  f;
  for (int k = 0; k < 5; k++) {
    u.p[k] = vars_all[k];
    f.p[k] = flux_all[k];
  }
  int blocks = (nf + nb + 255) / 256;
  flux_faces_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(nf, nb, ranks, indices, nbr, normals, areas, u, f,
                                                                 speed);
  return cudaGetLastError();
}

template <typename T>
static int rk3_stage_impl(int stage, int64_t n, int nvar, const T* const* prev, const T* const* in, T* const* out,
                          T* const* flux, const T* vol, int cells_per_vol, T dt, void* stream) {
  if (stage < 1 || stage > 3 || nvar < 1 || nvar > 8 || n < 0 || !prev || !out || !flux || !vol ||
      cells_per_vol < 1 || (stage > 1 && !in))
    return cudaErrorInvalidValue;
  if (n == 0) return cudaSuccess;
  PtrsC<T, 8> p{}, s{};
  Ptrs<T, 8>  o{}, f{};
  for (int k = 0; k < nvar; k++) {
    p.p[k] = prev[k];
    s.p[k] = stage > 1 ? in[k] : nullptr;
    o.p[k] = out[k];
    f.p[k] = flux[k];
  }
  unsigned     blocks = (unsigned)((n + 255) / 256);
  cudaStream_t st     = (cudaStream_t)stream;
  if (stage == 1)
    rk3_stage_kernel<T, 1><<<blocks, 256, 0, st>>>(n, nvar, p, s, o, f, vol, cells_per_vol, dt);
  else if (stage == 2)
    rk3_stage_kernel<T, 2><<<blocks, 256, 0, st>>>(n, nvar, p, s, o, f, vol, cells_per_vol, dt);
  else
    rk3_stage_kernel<T, 3><<<blocks, 256, 0, st>>>(n, nvar, p, s, o, f, vol, cells_per_vol, dt);
  return cudaGetLastError();
}

template <typename T>
static int max_speed_impl(const T* speed, int64_t n, T* out_dev, void* stream) {
  if (!out_dev || n < 0 || (n > 0 && !speed)) return cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t  e  = cudaMemsetAsync(out_dev, 0, sizeof(T), st);
  if (e != cudaSuccess) return e;
  if (n == 0) return cudaSuccess;
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  max_speed_kernel<T><<<blocks, 256, 0, st>>>(speed, n, out_dev);
  return cudaGetLastError();
}

// ============================================================================================================
// C ABI
// ============================================================================================================
extern "C" {

int t8b200_version(void) { return 100; }

int t8b200_flux_faces_f32(int32_t nf, int32_t nb, const int32_t* ranks, const int32_t* indices, const int32_t* nbr,
                          const float* normals, const float* areas, const float* const* const* vars_all,
                          float* const* const* flux_all, float* speed, void* stream) {
  return flux_faces_impl<float>(nf, nb, ranks, indices, nbr, normals, areas, vars_all, flux_all, speed, stream);
}
int t8b200_flux_faces_f64(int32_t nf, int32_t nb, const int32_t* ranks, const int32_t* indices, const int32_t* nbr,
                          const double* normals, const double* areas, const double* const* const* vars_all,
                          double* const* const* flux_all, double* speed, void* stream) {
  return flux_faces_impl<double>(nf, nb, ranks, indices, nbr, normals, areas, vars_all, flux_all, speed, stream);
}
int t8b200_rk3_stage_f32(int stage, int64_t n, int nvar, const float* const* prev, const float* const* in,
                         float* const* out, float* const* flux, const float* vol, int cpv, float dt, void* stream) {
  return rk3_stage_impl<float>(stage, n, nvar, prev, in, out, flux, vol, cpv, dt, stream);
}
int t8b200_rk3_stage_f64(int stage, int64_t n, int nvar, const double* const* prev, const double* const* in,
                         double* const* out, double* const* flux, const double* vol, int cpv, double dt,
                         void* stream) {
  return rk3_stage_impl<double>(stage, n, nvar, prev, in, out, flux, vol, cpv, dt, stream);
}
int t8b200_max_speed_f32(const float* speed, int64_t n, float* out_dev, void* stream) {
  return max_speed_impl<float>(speed, n, out_dev, stream);
}
int t8b200_max_speed_f64(const double* speed, int64_t n, double* out_dev, void* stream) {
  return max_speed_impl<double>(speed, n, out_dev, stream);
}

}  // extern "C"
