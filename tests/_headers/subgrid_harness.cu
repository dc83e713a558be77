// TEST-ONLY: C-ABI harness around the PRODUCT's t8gpu::SubgridMeshManager (include/t8gpu/mesh/subgrid_mesh_manager.h)
// over the t8mini stand-in for t8code.  Mirrors how examples/subgrid/solver.inl drives the reference's manager.
#include <t8gpu/mesh/subgrid_mesh_manager.h>
#include <t8gpu/timestepping/ssp_runge_kutta.h>

enum VariableList { Rho, Rho_v1, Rho_v2, Rho_v3, Rho_e, nb_variables };
enum StepList { Step0, Step1, Step2, Step3, Fluxes, nb_steps };

using S3 = t8gpu::Subgrid<4, 4, 4>;
using S2 = t8gpu::Subgrid<4, 4>;
using M3 = t8gpu::SubgridMeshManager<VariableList, StepList, S3>;
using M2 = t8gpu::SubgridMeshManager<VariableList, StepList, S2>;
using T  = M3::float_type;

struct Harness {
  int      dim;
  M3*      m3 = nullptr;
  M2*      m2 = nullptr;
  StepList next = Step0, prev = Step3;
};
#define DISPATCH(h, expr) ((h)->dim == 3 ? [&](auto* m) { return expr; }((h)->m3) : [&](auto* m) { return expr; }((h)->m2))

template <typename M>
static void get_conn(M* m, int32_t* ranks, int32_t* indices, int32_t* nbr, T* normals, T* areas, int32_t* ld,
                     int32_t* off, T* volumes) {
  auto          c = m->get_connectivity_information();
  const int     d = M::dim;
  const int64_t n = m->get_num_local_elements() + m->get_num_ghost_elements();
  const int64_t nf = c.get_num_local_faces(), nb = c.get_num_local_boundary_faces();
  cudaMemcpy(ranks, c.ranks(), sizeof(int32_t) * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(indices, c.indices(), sizeof(int32_t) * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(nbr, c.face_neighbors(), sizeof(int32_t) * (2 * nf + nb), cudaMemcpyDeviceToHost);
  cudaMemcpy(normals, c.face_normals(), sizeof(T) * d * (nf + nb), cudaMemcpyDeviceToHost);
  cudaMemcpy(areas, c.face_surfaces(), sizeof(T) * (nf + nb), cudaMemcpyDeviceToHost);
  cudaMemcpy(ld, c.face_level_difference(), sizeof(int32_t) * nf, cudaMemcpyDeviceToHost);
  cudaMemcpy(off, c.face_neighbor_offset(), sizeof(int32_t) * d * nf, cudaMemcpyDeviceToHost);
  cudaMemcpy(volumes, m->get_own_volume(), sizeof(T) * m->get_num_local_elements(), cudaMemcpyDeviceToHost);
}
template <typename M>
static void copy_state(M* m, StepList step, T* host, bool to_device) {
  const size_t n   = (size_t)m->get_num_local_elements() * M::dim == 0 ? 0 : (size_t)m->get_num_local_elements();
  const size_t len = n * (M::dim == 3 ? 64 : 16);
  auto         own = m->get_own_variables(step);
  cudaDeviceSynchronize();
  for (int k = 0; k < 5; k++) {
    if (to_device) cudaMemcpy(own.data()[k], host + k * len, sizeof(T) * len, cudaMemcpyHostToDevice);
    else cudaMemcpy(host + k * len, own.data()[k], sizeof(T) * len, cudaMemcpyDeviceToHost);
  }
}
template <typename M>
static void iterate(M* m, Harness* h, T dt, int nsteps) {
  for (int s = 0; s < nsteps; s++) {
    std::swap(h->next, h->prev);
    m->fused_stage(1, h->prev, h->prev, Step1, dt);
    m->fused_stage(2, Step1, h->prev, Step2, dt);
    m->fused_stage(3, Step2, h->prev, h->next, dt);
  }
  cudaDeviceSynchronize();
}
template <typename M>
static void criteria(M* m, Harness* h, T* out) {
  int n = m->get_num_local_elements();
  thrust::device_vector<T> c(n);
  m->refinement_criteria(h->next, Rho, thrust::raw_pointer_cast(c.data()));
  cudaMemcpy(out, thrust::raw_pointer_cast(c.data()), sizeof(T) * n, cudaMemcpyDeviceToHost);
}
template <typename M>
static void adapt(M* m, Harness* h, const T* crit) {
  int n = m->get_num_local_elements();
  thrust::host_vector<T> c(crit, crit + n);
  m->adapt(c, h->next);            // SubgridCompressibleEulerSolver::adapt, solver.inl:342-344
  m->partition(h->next);
  m->compute_connectivity_information();
  cudaDeviceSynchronize();
}
template <typename M>
static M* make(int dim, int level, int periodic) {
  t8_scheme_cxx_t* scheme = t8_scheme_new_default_cxx();
  t8_cmesh_t       cmesh  = t8mini_cmesh_new_cube(dim, periodic);
  t8_forest_t      forest = t8_forest_new_uniform(cmesh, scheme, level, true, sc_MPI_COMM_WORLD);
  M*               m      = new M(sc_MPI_COMM_WORLD, scheme, cmesh, forest);
  m->initialize_variables([](t8gpu::MemoryAccessorOwn<VariableList>& u, t8_forest_t f, t8_locidx_t tree,
                             t8_element_t const* element, t8_locidx_t e) {
    double c[3];
    t8_forest_element_centroid(f, tree, element, c);
    auto [rho, m1, m2, m3, en] = u.get(Rho, Rho_v1, Rho_v2, Rho_v3, Rho_e);
    rho[e] = T(1.0 + 0.25 * c[0]); m1[e] = T(0.1); m2[e] = T(-0.05 * c[1]); m3[e] = T(0.0); en[e] = T(2.5 / 0.4 + 0.1);
  });
  return m;
}

extern "C" {
int   sh_float_size() { return (int)sizeof(T); }
void* sh_create(int dim, int level, int periodic) {
  auto* h = new Harness{dim};
  if (dim == 3) h->m3 = make<M3>(dim, level, periodic); else h->m2 = make<M2>(dim, level, periodic);
  return h;
}
void sh_destroy(void* p) {
  auto* h = static_cast<Harness*>(p);
  delete h->m3;
  delete h->m2;
  delete h;
}
void sh_counts(void* p, int64_t out[4]) {
  auto* h = static_cast<Harness*>(p);
  out[0] = DISPATCH(h, m->get_num_local_elements()); out[1] = DISPATCH(h, m->get_num_ghost_elements());
  out[2] = DISPATCH(h, m->get_num_local_faces());    out[3] = DISPATCH(h, m->get_num_local_boundary_faces());
}
void sh_get_connectivity(void* p, int32_t* ranks, int32_t* indices, int32_t* nbr, T* normals, T* areas, int32_t* ld,
                         int32_t* off, T* volumes) {
  auto* h = static_cast<Harness*>(p);
  DISPATCH(h, get_conn(m, ranks, indices, nbr, normals, areas, ld, off, volumes));
}
void sh_set_state(void* p, T* u) { auto* h = static_cast<Harness*>(p); DISPATCH(h, copy_state(m, h->next, u, true)); }
void sh_get_state(void* p, T* u) { auto* h = static_cast<Harness*>(p); DISPATCH(h, copy_state(m, h->next, u, false)); }
void sh_iterate(void* p, double dt, int n) { auto* h = static_cast<Harness*>(p); DISPATCH(h, iterate(m, h, (T)dt, n)); }
void sh_criteria(void* p, T* out) { auto* h = static_cast<Harness*>(p); DISPATCH(h, criteria(m, h, out)); }
void sh_adapt(void* p, const T* crit) { auto* h = static_cast<Harness*>(p); DISPATCH(h, adapt(m, h, crit)); }
int  sh_last_cuda_error() { return (int)cudaGetLastError(); }
}
