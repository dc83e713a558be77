// TEST INFRASTRUCTURE ONLY -- C-ABI harness around the reference's OWN unstructured solver
// (examples/compressible_euler/{solver,kernels}.cu compiled unmodified from /root/reference, linked against t8mini).
// Built by oracle/ref_build.py into oracle/_ref/libref_uns_{f32,f64}.so.  Used by tests (golden generation, parity) and
// by bench.py --impl reference.  `private` is opened up in THIS translation unit only, to read the mesh manager's
// device arrays and to overwrite the initial state; the reference's own objects are compiled as they are.
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
#include <thrust/reduce.h>

#define private public
#define protected public
#include "solver.h"
#include "kernels.h"
#undef private
#undef protected

using namespace t8gpu;
using T = CompressibleEulerSolver::float_type;

template <typename V>
static void d2h(V const& dv, void* out) {
  if (dv.size())
    cudaMemcpy(out, thrust::raw_pointer_cast(dv.data()), dv.size() * sizeof(typename V::value_type),
               cudaMemcpyDeviceToHost);
}

__global__ static void crit_div(T const* fluxes_rho, T const* volume, T* criteria, int n) {
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  criteria[i] = fluxes_rho[i] / cbrt(volume[i]);   // solver.cu:243
}

// The reference never initialises its flux accumulators after a (re)allocation; inside a long test
// process that memory is recycled.  The harness clears the array where the reference relies on it being zero.
static void zero_fluxes(CompressibleEulerSolver* s) {
  int n = s->m_mesh_manager.get_num_local_elements();
  for (int k = 0; k < 5; k++)
    cudaMemset(s->m_mesh_manager.get_own_variable(Fluxes, static_cast<VariableList>(k)), 0, sizeof(T) * n);
}

extern "C" {

int ref_float_size() { return (int)sizeof(T); }

void* ref_create(int dim, int level, int periodic) {
  t8_scheme_cxx_t* scheme = t8_scheme_new_default_cxx();
  t8_cmesh_t       cmesh  = t8mini_cmesh_new_cube(dim, periodic);
  t8_forest_t      forest = t8_forest_new_uniform(cmesh, scheme, level, true, sc_MPI_COMM_WORLD);
  auto* s = new CompressibleEulerSolver(sc_MPI_COMM_WORLD, scheme, cmesh, forest);
  zero_fluxes(s);
  return s;
}
void ref_destroy(void* h) { delete static_cast<CompressibleEulerSolver*>(h); }

// out: n_elements, n_ghosts, n_faces, n_boundary_faces
void ref_counts(void* h, int64_t out[4]) {
  auto* s = static_cast<CompressibleEulerSolver*>(h);
  out[0]  = s->m_mesh_manager.get_num_local_elements();
  out[1]  = s->m_mesh_manager.get_num_ghost_elements();
  out[2]  = s->m_mesh_manager.get_num_local_faces();
  out[3]  = s->m_mesh_manager.get_num_local_boundary_faces();
}

// the arrays MeshManager::compute_connectivity_information uploaded (mesh_manager.inl:332-481) + volumes + levels
void ref_get_connectivity(void* h, int32_t* ranks, int32_t* indices, int32_t* nbr, T* normals, T* areas, T* volumes) {
  auto* s = static_cast<CompressibleEulerSolver*>(h);
  auto& m = s->m_mesh_manager;
  d2h(m.m_device_ranks, ranks);
  d2h(m.m_device_indices, indices);
  d2h(m.m_device_face_neighbors, nbr);
  d2h(m.m_device_face_normals, normals);
  d2h(m.m_device_face_area, areas);
  cudaMemcpy(volumes, m.get_own_volume(), sizeof(T) * m.get_num_local_elements(), cudaMemcpyDeviceToHost);
}

void ref_set_state(void* h, const T* u) {
  auto* s = static_cast<CompressibleEulerSolver*>(h);
  int   n = s->m_mesh_manager.get_num_local_elements();
  for (int k = 0; k < 5; k++)
    cudaMemcpy(s->m_mesh_manager.get_own_variable(s->next, static_cast<VariableList>(k)), u + (size_t)k * n,
               sizeof(T) * n, cudaMemcpyHostToDevice);
}
void ref_get_state(void* h, T* u) {
  auto* s = static_cast<CompressibleEulerSolver*>(h);
  int   n = s->m_mesh_manager.get_num_local_elements();
  cudaDeviceSynchronize();
  for (int k = 0; k < 5; k++)
    cudaMemcpy(u + (size_t)k * n, s->m_mesh_manager.get_own_variable(s->next, static_cast<VariableList>(k)),
               sizeof(T) * n, cudaMemcpyDeviceToHost);
}
void ref_iterate(void* h, double dt, int nsteps) {
  auto* s = static_cast<CompressibleEulerSolver*>(h);
  for (int i = 0; i < nsteps; i++) s->iterate(static_cast<T>(dt));
  cudaDeviceSynchronize();
}
double ref_compute_timestep(void* h) { return (double)static_cast<CompressibleEulerSolver*>(h)->compute_timestep(); }
void   ref_adapt(void* h) {
  static_cast<CompressibleEulerSolver*>(h)->adapt();
  zero_fluxes(static_cast<CompressibleEulerSolver*>(h));
  cudaDeviceSynchronize();
}

// milliseconds for `steps` calls of the reference's own iterate() (its cudaDeviceSynchronize()s included), CUDA events
double ref_time_steps(void* h, double dt, int warmup, int steps) {
  auto* s = static_cast<CompressibleEulerSolver*>(h);
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
// MeshManager::adapt with caller-supplied criteria (mesh_manager.inl:195-330) + connectivity rebuild, as
// CompressibleEulerSolver::adapt does after computing its own criteria (solver.cu:273-276)
void ref_mesh_adapt(void* h, const T* crit) {
  auto* s = static_cast<CompressibleEulerSolver*>(h);
  int   n = s->m_mesh_manager.get_num_local_elements();
  thrust::host_vector<T> c(crit, crit + n);
  s->m_mesh_manager.adapt(c, s->next);
  s->m_mesh_manager.compute_connectivity_information();
  s->m_device_face_speed_estimate.resize(s->m_mesh_manager.get_num_local_faces() +
                                         s->m_mesh_manager.get_num_local_boundary_faces());
  zero_fluxes(s);
  cudaDeviceSynchronize();
}
// Refinement criteria exactly as CompressibleEulerSolver::adapt computes them (solver.cu:246-271): the reference's own
// estimate_gradient kernel, then fluxes_rho / cbrt(volume) (compute_refinement_criteria is file-static in solver.cu, so
// its one line is restated in crit_div below), then the flux array is cleared again.
void ref_criteria(void* h, T* out) {
  auto* s = static_cast<CompressibleEulerSolver*>(h);
  auto& m = s->m_mesh_manager;
  int   n = m.get_num_local_elements(), nf = m.get_num_local_faces();
  if (nf > 0)
    t8gpu::estimate_gradient<<<(nf + 255) / 256, 256>>>(m.get_connectivity_information(), m.get_all_variables(s->next),
                                                       m.get_all_variables(Fluxes));
  cudaDeviceSynchronize();
  thrust::device_vector<T> c(n);
  crit_div<<<(n + 255) / 256, 256>>>(m.get_own_variable(Fluxes, Rho), m.get_own_volume(),
                                     thrust::raw_pointer_cast(c.data()), n);
  cudaMemcpy(out, thrust::raw_pointer_cast(c.data()), sizeof(T) * n, cudaMemcpyDeviceToHost);
  cudaMemset(m.get_own_variable(Fluxes, Rho), 0, sizeof(T) * n);
}
int ref_last_cuda_error() { return (int)cudaGetLastError(); }
// output path: CompressibleEulerSolver::save_conserved_variables_to_vtk (solver.cu:177-186) ->
// MeshManager::get_host_{scalar,vector}_variable + save_variables_to_vtk (mesh_manager.inl:515-623)
void ref_save_conserved(void* h, const char* prefix) {
  static_cast<CompressibleEulerSolver*>(h)->save_conserved_variables_to_vtk(prefix);
  cudaDeviceSynchronize();
}
}
