// Output path of the subgrid manager (SURVEY f-4): cell data of Subgrid<4,4,4> / Subgrid<4,4> elements from the
// manager's column-major order (x fastest inside an element) to the Morton order in which t8code enumerates the leaves
// of the forest refined log2(4) = 2 more times -- what t8_forest_write_vtk_ext needs as element data.
//
// Reference behaviour replaced (not translated): column_major_to_z_order<<<N, (4,4,4)>>> into a device temporary, a
// device -> host copy and, for float_type = float, a host loop that widens every value to double
// (t8gpu/mesh/subgrid_mesh_manager.inl:1007-1124).  Here one kernel permutes and widens, so the host receives the
// `double` array t8code takes in ONE copy: thread g writes out[g] (coalesced), reading the cell whose Morton index
// inside its element is g % S (the 64 / 16 cells of an element share two / one 128-byte lines either way).
#include <cuda_runtime.h>

#include "../../include/t8gpu_b200.h"

namespace {

template <int DIM>
__device__ __forceinline__ int flat_of_morton(int m) {
  if (DIM == 3) {   // m = i0 j0 k0 i1 j1 k1 (bit 0 first); flat = i + 4 j + 16 k
    const int i = (m & 1) | ((m >> 2) & 2), j = ((m >> 1) & 1) | ((m >> 3) & 2), k = ((m >> 2) & 1) | ((m >> 4) & 2);
    return i + 4 * j + 16 * k;
  }
  const int i = (m & 1) | ((m >> 1) & 2), j = ((m >> 1) & 1) | ((m >> 2) & 2);   // m = i0 j0 i1 j1; flat = i + 4 j
  return i + 4 * j;
}

template <typename T, int DIM>
__global__ void __launch_bounds__(256) z_order_kernel(int64_t n_cells, const T* __restrict__ from,
                                                      double* __restrict__ to) {
  constexpr int S = DIM == 3 ? 64 : 16;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_cells) return;
  const int64_t e = g / S;
  to[g] = (double)from[e * S + flat_of_morton<DIM>((int)(g % S))];
}

template <typename T>
int z_order_impl(int dim, int64_t n_elements, const T* from, double* to, void* stream) {
  if ((dim != 2 && dim != 3) || n_elements < 0) return cudaErrorInvalidValue;
  if (n_elements == 0) return 0;
  if (!from || !to || (const void*)from == (const void*)to) return cudaErrorInvalidValue;
  const int64_t  cells = n_elements * (dim == 3 ? 64 : 16);
  const unsigned grid  = (unsigned)((cells + 255) / 256);
  if (dim == 3) z_order_kernel<T, 3><<<grid, 256, 0, (cudaStream_t)stream>>>(cells, from, to);
  else z_order_kernel<T, 2><<<grid, 256, 0, (cudaStream_t)stream>>>(cells, from, to);
  return cudaGetLastError();
}

}  // namespace

extern "C" {
int t8b200_subgrid_z_order_f32(int dim, int64_t n_elements, const float* from, double* to, void* stream) {
  return z_order_impl<float>(dim, n_elements, from, to, stream);
}
int t8b200_subgrid_z_order_f64(int dim, int64_t n_elements, const double* from, double* to, void* stream) {
  return z_order_impl<double>(dim, n_elements, from, to, stream);
}
}
