// TEST-ONLY: C-ABI harness around the PRODUCT's header mirror (include/t8gpu/mesh/mesh_manager.h) running over the
// t8mini stand-in for t8code (oracle/ref_shim, oracle/miniforest.c).  Lets the parity tests drive
// t8gpu::MeshManager<VariableList, StepList, 3> exactly as the reference's example solver drives its own:
// construct from a forest, initialise, read the connectivity, step (here with the fused stage), adapt.
#include <t8gpu/mesh/mesh_manager.h>
#include <t8gpu/timestepping/ssp_runge_kutta.h>

#include <cstring>

enum VariableList { Rho, Rho_v1, Rho_v2, Rho_v3, Rho_e, nb_variables };
enum StepList { Step0, Step1, Step2, Step3, Fluxes, nb_steps };

using Mesh = t8gpu::MeshManager<VariableList, StepList, 3>;
using T    = Mesh::float_type;

struct Harness {
  Mesh*    mesh;
  StepList next = Step0, prev = Step3;
  T*       speed_max = nullptr;
};

extern "C" {
int   mh_float_size() { return (int)sizeof(T); }
void* mh_create(int dim, int level, int periodic) {
  t8_scheme_cxx_t* scheme = t8_scheme_new_default_cxx();
  t8_cmesh_t       cmesh  = t8mini_cmesh_new_cube(dim, periodic);
  t8_forest_t      forest = t8_forest_new_uniform(cmesh, scheme, level, true, sc_MPI_COMM_WORLD);
  auto*            h      = new Harness{new Mesh(sc_MPI_COMM_WORLD, scheme, cmesh, forest)};
  h->mesh->initialize_variables([](t8gpu::MemoryAccessorOwn<VariableList>& u, t8_forest_t f, t8_locidx_t tree,
                                   t8_element_t const* element, t8_locidx_t e) {
    double c[3];
    t8_forest_element_centroid(f, tree, element, c);
    auto [rho, m1, m2, m3, en] = u.get(Rho, Rho_v1, Rho_v2, Rho_v3, Rho_e);
    rho[e] = T(1.0 + 0.25 * c[0]); m1[e] = T(0.1); m2[e] = T(-0.05 * c[1]); m3[e] = T(0.0); en[e] = T(2.5 / 0.4 + 0.1);
  });
  cudaMalloc(&h->speed_max, sizeof(T));
  return h;
}
void mh_destroy(void* p) {
  auto* h = static_cast<Harness*>(p);
  cudaFree(h->speed_max);
  delete h->mesh;
  delete h;
}
void mh_counts(void* p, int64_t out[4]) {
  auto* m = static_cast<Harness*>(p)->mesh;
  out[0] = m->get_num_local_elements(); out[1] = m->get_num_ghost_elements();
  out[2] = m->get_num_local_faces();    out[3] = m->get_num_local_boundary_faces();
}
void mh_get_connectivity(void* p, int32_t* ranks, int32_t* indices, int32_t* nbr, T* normals, T* areas, T* volumes) {
  auto* m = static_cast<Harness*>(p)->mesh;
  auto  c = m->get_connectivity_information();
  const int64_t n = m->get_num_local_elements() + m->get_num_ghost_elements();
  const int64_t nf = c.get_num_local_faces(), nb = c.get_num_local_boundary_faces();
  cudaMemcpy(ranks, c.ranks(), sizeof(int32_t) * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(indices, c.indices(), sizeof(int32_t) * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(nbr, c.face_neighbors(), sizeof(int32_t) * (2 * nf + nb), cudaMemcpyDeviceToHost);
  cudaMemcpy(normals, c.face_normals(), sizeof(T) * 3 * (nf + nb), cudaMemcpyDeviceToHost);
  cudaMemcpy(areas, c.face_surfaces(), sizeof(T) * (nf + nb), cudaMemcpyDeviceToHost);
  cudaMemcpy(volumes, m->get_own_volume(), sizeof(T) * m->get_num_local_elements(), cudaMemcpyDeviceToHost);
}
void mh_set_state(void* p, const T* u) {
  auto* h = static_cast<Harness*>(p);
  int   n = h->mesh->get_num_local_elements();
  for (int k = 0; k < 5; k++)
    cudaMemcpy(h->mesh->get_own_variable(h->next, static_cast<VariableList>(k)), u + (size_t)k * n, sizeof(T) * n,
               cudaMemcpyHostToDevice);
}
void mh_get_state(void* p, T* u) {
  auto* h = static_cast<Harness*>(p);
  int   n = h->mesh->get_num_local_elements();
  cudaDeviceSynchronize();
  for (int k = 0; k < 5; k++)
    cudaMemcpy(u + (size_t)k * n, h->mesh->get_own_variable(h->next, static_cast<VariableList>(k)), sizeof(T) * n,
               cudaMemcpyDeviceToHost);
}
// CompressibleEulerSolver::iterate (solver.cu:75-175) with the fused stages
void mh_iterate(void* p, double dt, int nsteps) {
  auto* h = static_cast<Harness*>(p);
  for (int s = 0; s < nsteps; s++) {
    std::swap(h->next, h->prev);
    h->mesh->fused_stage(1, h->prev, h->prev, Step1, (T)dt);
    h->mesh->fused_stage(2, Step1, h->prev, Step2, (T)dt);
    h->mesh->fused_stage(3, Step2, h->prev, h->next, (T)dt, h->speed_max);
  }
  cudaDeviceSynchronize();
}
// the same step with the reference-shaped launches on the accessors (what an unmodified user solver does)
void mh_iterate_unfused(void* p, double dt, int nsteps) {
  auto* h = static_cast<Harness*>(p);
  auto& m = *h->mesh;
  auto  flux = [&](StepList in) {
    auto c = m.get_connectivity_information();
    if (sizeof(T) == 8)
      t8b200_flux_faces_f64(c.get_num_local_faces(), c.get_num_local_boundary_faces(), c.ranks(), c.indices(),
                            c.face_neighbors(), (const double*)c.face_normals(), (const double*)c.face_surfaces(),
                            (const double* const* const*)m.get_all_variables(in).data(),
                            (double* const* const*)m.get_all_variables(Fluxes).data(), nullptr, nullptr);
    else
      t8b200_flux_faces_f32(c.get_num_local_faces(), c.get_num_local_boundary_faces(), c.ranks(), c.indices(),
                            c.face_neighbors(), (const float*)c.face_normals(), (const float*)c.face_surfaces(),
                            (const float* const* const*)m.get_all_variables(in).data(),
                            (float* const* const*)m.get_all_variables(Fluxes).data(), nullptr, nullptr);
  };
  const int n = m.get_num_local_elements(), blocks = (n + 255) / 256;
  for (int s = 0; s < nsteps; s++) {
    std::swap(h->next, h->prev);
    flux(h->prev);
    t8gpu::timestepping::SSP_3RK_step1<VariableList><<<blocks, 256>>>(
        m.get_own_variables(h->prev), m.get_own_variables(Step1), m.get_own_variables(Fluxes), m.get_own_volume(), (T)dt, n);
    flux(Step1);
    t8gpu::timestepping::SSP_3RK_step2<VariableList><<<blocks, 256>>>(
        m.get_own_variables(h->prev), m.get_own_variables(Step1), m.get_own_variables(Step2), m.get_own_variables(Fluxes),
        m.get_own_volume(), (T)dt, n);
    flux(Step2);
    t8gpu::timestepping::SSP_3RK_step3<VariableList><<<blocks, 256>>>(
        m.get_own_variables(h->prev), m.get_own_variables(Step2), m.get_own_variables(h->next),
        m.get_own_variables(Fluxes), m.get_own_volume(), (T)dt, n);
  }
  cudaDeviceSynchronize();
}
double mh_speed_max(void* p) {
  T v = 0;
  cudaMemcpy(&v, static_cast<Harness*>(p)->speed_max, sizeof(T), cudaMemcpyDeviceToHost);
  return (double)v;
}
void mh_criteria(void* p, T* out) {
  auto* h = static_cast<Harness*>(p);
  int   n = h->mesh->get_num_local_elements();
  thrust::device_vector<T> c(n);
  h->mesh->gradient_criteria(h->next, Rho, thrust::raw_pointer_cast(c.data()));
  cudaMemcpy(out, thrust::raw_pointer_cast(c.data()), sizeof(T) * n, cudaMemcpyDeviceToHost);
}
// CompressibleEulerSolver::adapt after the criteria (solver.cu:273-276)
void mh_adapt(void* p, const T* crit) {
  auto* h = static_cast<Harness*>(p);
  int   n = h->mesh->get_num_local_elements();
  thrust::host_vector<T> c(crit, crit + n);
  h->mesh->adapt(c, h->next);
  h->mesh->partition(h->next);
  h->mesh->compute_connectivity_information();
  cudaDeviceSynchronize();
}
int mh_last_cuda_error() { return (int)cudaGetLastError(); }
}
