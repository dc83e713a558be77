// TEST INFRASTRUCTURE ONLY -- C-ABI harness around the reference's OWN subgrid solver
// (examples/subgrid/{solver,kernels}_{2d,3d}.cu compiled unmodified from /root/reference, linked against t8mini).
// Built by oracle/ref_build.py into oracle/_ref/libref_sg_{f32,f64}.so.  See harness_uns.cu.
#include <array>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include <cuda_runtime.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#define private public
#define protected public
#include "solver.h"
#include "kernels.h"
#undef private
#undef protected

using S3 = SubgridCompressibleEulerSolver<t8gpu::Subgrid<4, 4, 4>>;
using S2 = SubgridCompressibleEulerSolver<t8gpu::Subgrid<4, 4>>;
using T  = S3::float_type;

struct H {
  int dim;
  S3* s3 = nullptr;
  S2* s2 = nullptr;
};

template <typename V>
static void d2h(V const& dv, void* out) {
  if (dv.size())
    cudaMemcpy(out, thrust::raw_pointer_cast(dv.data()), dv.size() * sizeof(typename V::value_type),
               cudaMemcpyDeviceToHost);
}

#define DISPATCH(h, expr) (static_cast<H*>(h)->dim == 3 ? [&](auto* s) { return expr; }(static_cast<H*>(h)->s3) \
                                                        : [&](auto* s) { return expr; }(static_cast<H*>(h)->s2))

template <typename S>
static void counts(S* s, int64_t out[4]) {
  out[0] = s->m_mesh_manager.get_num_local_elements();
  out[1] = s->m_mesh_manager.get_num_ghost_elements();
  out[2] = s->m_mesh_manager.get_num_local_faces();
  out[3] = s->m_mesh_manager.get_num_local_boundary_faces();
}
template <typename S>
static void conn(S* s, int32_t* ranks, int32_t* indices, int32_t* nbr, T* normals, T* areas, int32_t* ld, int32_t* off,
                 T* volumes) {
  auto& m = s->m_mesh_manager;
  d2h(m.m_device_ranks, ranks);
  d2h(m.m_device_indices, indices);
  d2h(m.m_device_face_neighbors, nbr);
  d2h(m.m_device_face_normals, normals);
  d2h(m.m_device_face_area, areas);
  d2h(m.m_device_face_level_difference, ld);
  d2h(m.m_device_face_neighbor_offset, off);
  cudaMemcpy(volumes, m.get_own_volume(), sizeof(T) * m.get_num_local_elements(), cudaMemcpyDeviceToHost);
}
template <typename S>
static void set_state(S* s, const T* u) {
  size_t n = (size_t)s->m_mesh_manager.get_num_local_elements() * S::subgrid_type::size;
  for (int k = 0; k < 5; k++)
    cudaMemcpy(static_cast<T*>(s->m_mesh_manager.get_own_variable(s->next, static_cast<VariableList>(k))), u + k * n, sizeof(T) * n,
               cudaMemcpyHostToDevice);
}
template <typename S>
static void get_state(S* s, T* u) {
  size_t n = (size_t)s->m_mesh_manager.get_num_local_elements() * S::subgrid_type::size;
  cudaDeviceSynchronize();
  for (int k = 0; k < 5; k++)
    cudaMemcpy(u + k * n, static_cast<T*>(s->m_mesh_manager.get_own_variable(s->next, static_cast<VariableList>(k))), sizeof(T) * n,
               cudaMemcpyDeviceToHost);
}
template <typename S>
static double time_steps(S* s, double dt, int warmup, int steps) {
  for (int i = 0; i < warmup; i++) s->iterate(static_cast<T>(dt));
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  for (int i = 0; i < steps; i++) s->iterate(static_cast<T>(dt));
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return (double)ms;
}

// The reference never initialises its flux accumulators: a new allocation (constructor, or a resize that had to grow,
// SURVEY App. D-14) holds whatever the allocator hands back, and iterate() accumulates into it.  On a fresh device that
// is zero; inside a long test process it is recycled memory.  The harness clears the array where the reference relies on
// it being zero -- nothing else of the reference's state is touched.
template <typename S>
static void zero_fluxes(S* s) {
  size_t n = (size_t)s->m_mesh_manager.get_num_local_elements() * S::subgrid_type::size;
  for (int k = 0; k < 5; k++)
    cudaMemset(static_cast<T*>(s->m_mesh_manager.get_own_variable(Fluxes, static_cast<VariableList>(k))), 0, sizeof(T) * n);
}

template <typename S>
static void mesh_adapt(S* s, const T* crit) {
  int n = s->m_mesh_manager.get_num_local_elements();
  thrust::host_vector<T> c(crit, crit + n);
  s->m_mesh_manager.adapt(c, s->next);      // subgrid_mesh_manager.inl:427-558
  s->m_mesh_manager.partition(s->next);     // :1282-1369 (as SubgridCompressibleEulerSolver::adapt, solver.inl:342-344)
  s->m_mesh_manager.compute_connectivity_information();
  zero_fluxes(s);
  cudaDeviceSynchronize();
}

// the reference's own compute_refinement_criteria kernel, launched as SubgridCompressibleEulerSolver::adapt does
// (solver.inl:331-339)
template <typename S>
static void criteria(S* s, T* out) {
  using SubgridType = typename S::subgrid_type;
  int n = s->m_mesh_manager.get_num_local_elements();
  thrust::device_vector<T> c(n);
  compute_refinement_criteria<SubgridType><<<(n + 255) / 256, 256>>>(
      s->m_mesh_manager.get_own_variable(s->next, Rho), thrust::raw_pointer_cast(c.data()),
      s->m_mesh_manager.get_own_volume(), n);
  cudaDeviceSynchronize();
  cudaMemcpy(out, thrust::raw_pointer_cast(c.data()), sizeof(T) * n, cudaMemcpyDeviceToHost);
}

extern "C" {
int ref_float_size() { return (int)sizeof(T); }

void* ref_create(int dim, int level, int periodic) {
  t8_scheme_cxx_t* scheme = t8_scheme_new_default_cxx();
  t8_cmesh_t       cmesh  = t8mini_cmesh_new_cube(dim, periodic);
  t8_forest_t      forest = t8_forest_new_uniform(cmesh, scheme, level, true, sc_MPI_COMM_WORLD);
  H*               h      = new H{dim};
  if (dim == 3)
    h->s3 = new S3(sc_MPI_COMM_WORLD, scheme, cmesh, forest);
  else
    h->s2 = new S2(sc_MPI_COMM_WORLD, scheme, cmesh, forest);
  DISPATCH(h, zero_fluxes(s));
  return h;
}
void ref_destroy(void* h) {
  H* p = static_cast<H*>(h);
  delete p->s3;
  delete p->s2;
  delete p;
}
void ref_counts(void* h, int64_t out[4]) { DISPATCH(h, counts(s, out)); }
void ref_get_connectivity(void* h, int32_t* ranks, int32_t* indices, int32_t* nbr, T* normals, T* areas, int32_t* ld,
                          int32_t* off, T* volumes) {
  DISPATCH(h, conn(s, ranks, indices, nbr, normals, areas, ld, off, volumes));
}
void ref_set_state(void* h, const T* u) { DISPATCH(h, set_state(s, u)); }
void ref_get_state(void* h, T* u) { DISPATCH(h, get_state(s, u)); }
void ref_iterate(void* h, double dt, int nsteps) {
  for (int i = 0; i < nsteps; i++) DISPATCH(h, s->iterate(static_cast<T>(dt)));
  cudaDeviceSynchronize();
}
void   ref_adapt(void* h) { DISPATCH(h, s->adapt()); DISPATCH(h, zero_fluxes(s)); cudaDeviceSynchronize(); }
double ref_time_steps(void* h, double dt, int warmup, int steps) { return DISPATCH(h, time_steps(s, dt, warmup, steps)); }
void   ref_mesh_adapt(void* h, const T* crit) { DISPATCH(h, mesh_adapt(s, crit)); }
void   ref_criteria(void* h, T* out) { DISPATCH(h, criteria(s, out)); }
int    ref_last_cuda_error() { return (int)cudaGetLastError(); }
// output path: SubgridCompressibleEulerSolver::save_density_to_vtk / save_mesh_to_vtk (solver.inl:268-279) ->
// SubgridMeshManager::save_variable_to_vtk / save_mesh_to_vtk (subgrid_mesh_manager.inl:1051-1142)
void ref_save_density(void* h, const char* prefix) { DISPATCH(h, s->save_density_to_vtk(prefix)); cudaDeviceSynchronize(); }
void ref_save_mesh(void* h, const char* prefix) { DISPATCH(h, s->save_mesh_to_vtk(prefix)); }
}
