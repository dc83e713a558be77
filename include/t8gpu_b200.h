/* t8gpu_b200 -- C ABI of the B200-native finite-volume solver core.
 *
 * This is the drop-in boundary for t8gpu's hot path (per-face numerical flux + SSP-RK3 stage update + CFL
 * wave-speed reduction).  The reference has no FFI for this path: user solvers launch `__global__` kernels on
 * accessor PODs obtained from t8gpu::MeshManager / MemoryManager.  The entry points below are what the header-only
 * template shim in include/t8gpu/ (same class and member names as the reference) binds to; each comment cites the
 * reference interface it replaces (paths relative to the t8gpu repository).
 *
 * Conventions
 *   - every pointer documented "device" must be dereferenceable by the current CUDA device (own or peer-mapped),
 *     "host" pointers by the CPU.  No C++ or torch types cross this boundary.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous unless stated.
 *   - return value: 0 on success, otherwise a cudaError_t value (cudaErrorInvalidValue for bad arguments).
 *     The C++ shim feeds it to T8GPU_CUDA_CHECK_ERROR to keep the reference's abort-on-error behaviour
 *     (t8gpu/utils/cuda.h:7-15).
 *   - precision suffix: _f32 / _f64 = variable_traits<...>::float_type (t8gpu/memory/memory_manager.h:29).
 *   - five conserved variables in the order of the examples' VariableList {Rho, Rho_v1, Rho_v2, Rho_v3, Rho_e}
 *     (examples/compressible_euler/solver.h, examples/subgrid/solver.h:12-20).
 */
#ifndef T8GPU_B200_H
#define T8GPU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T8B200_NVAR 5

/* ABI / build identification: returns 100 * major + minor. */
int t8b200_version(void);

/* -------------------------------------------------------------------------------------------------------------
 * 1. Reference-shaped kernels (same inputs, same outputs, same side effects as the reference's launches)
 * ------------------------------------------------------------------------------------------------------------- */

/* Replaces kepes_compute_fluxes<<<>>> + reflective_boundary_condition<<<>>>
 *   (examples/compressible_euler/kernels.cu:135-309, :311-469; launched at solver.cu:81-96).
 * Connectivity arrays are exactly those behind MeshConnectivityAccessor<float_type,3>
 *   (t8gpu/mesh/mesh_manager.h:159-166): ranks/indices[n_local+n_ghost], face_neighbors[2*nf+nb],
 *   face_normals[3*(nf+nb)] interleaved, face_surfaces[nf+nb].
 * vars_all / flux_all: HOST arrays of 5 DEVICE tables; table k holds one device pointer per rank
 *   (what MemoryAccessorAll<VariableList>::get(k) returns, memory_manager.h:240-246).
 * Accumulates -F into the left and +F into the right element with atomics, writes speed[f] = |uHat|+aHat for
 * f in [0, nf+nb).  speed may be NULL. */
int t8b200_flux_faces_f32(int32_t nf, int32_t nb, const int32_t* ranks, const int32_t* indices,
                          const int32_t* face_neighbors, const float* face_normals, const float* face_surfaces,
                          const float* const* const* vars_all, float* const* const* flux_all, float* speed,
                          void* stream);
int t8b200_flux_faces_f64(int32_t nf, int32_t nb, const int32_t* ranks, const int32_t* indices,
                          const int32_t* face_neighbors, const double* face_normals, const double* face_surfaces,
                          const double* const* const* vars_all, double* const* const* flux_all, double* speed,
                          void* stream);

/* Replaces timestepping::SSP_3RK_step{1,2,3}<VariableType> and timestepping::subgrid::SSP_3RK_step{1,2,3}
 *   (t8gpu/timestepping/ssp_runge_kutta.inl:30-99, :101-221).  stage in {1,2,3}.
 * prev/in/out/flux: HOST arrays of nvar DEVICE pointers (MemoryAccessorOwn<VariableType>::get(k)); `in` is the
 * previous stage (ignored for stage 1).  n = number of cells; volume of cell i = vol[i / cells_per_vol] / cells_per_vol
 * (cells_per_vol = 1 for elements, Subgrid::size for subgrids, ssp_runge_kutta.inl:116).  Zeroes flux. */
int t8b200_rk3_stage_f32(int stage, int64_t n, int nvar, const float* const* prev, const float* const* in,
                         float* const* out, float* const* flux, const float* vol, int cells_per_vol, float dt,
                         void* stream);
int t8b200_rk3_stage_f64(int stage, int64_t n, int nvar, const double* const* prev, const double* const* in,
                         double* const* out, double* const* flux, const double* vol, int cells_per_vol, double dt,
                         void* stream);

/* Replaces thrust::reduce(speed_estimates, 0, maximum) (examples/compressible_euler/solver.cu:214-217):
 * *out_dev = max(0, max_i speed[i]).  out_dev is a device scalar. */
int t8b200_max_speed_f32(const float* speed, int64_t n, float* out_dev, void* stream);
int t8b200_max_speed_f64(const double* speed, int64_t n, double* out_dev, void* stream);

/* Subgrid path (Subgrid<4,4,4> when dim = 3, Subgrid<4,4> when dim = 2; cell (e,i,j,k) at e*S + i + 4j + 16k,
 * t8gpu/memory/subgrid_memory_manager.h:35-135).
 * Replaces compute_inner_fluxes<<<N,(4,4,4)>>> (examples/subgrid/kernels.inl:335-662, launched at solver.inl:166):
 * fluxes between the cells of one element, accumulated non-atomically into this rank's flux arrays.
 * vars / flux: HOST arrays of 5 DEVICE pointers (SubgridMemoryAccessorOwn::get(k)); vol: device, per element. */
int t8b200_subgrid_inner_flux_f32(int dim, int64_t n_elements, const float* vol, const float* const* vars,
                                  float* const* flux, void* stream);
int t8b200_subgrid_inner_flux_f64(int dim, int64_t n_elements, const double* vol, const double* const* vars,
                                  double* const* flux, void* stream);
/* Replaces compute_outer_fluxes<<<Faces,(4,4)>>> (kernels.inl:664-911, solver.inl:181): fluxes across the faces
 * between elements, 2:1 hanging faces through level_difference / offset, atomics into both sides (possibly on
 * another rank).  Arrays are those behind SubgridMeshConnectivityAccessor (t8gpu/mesh/subgrid_mesh_manager.h:29-216):
 * normals have `dim` components per face, offsets `dim` per face.  vars_all / flux_all as in t8b200_flux_faces. */
int t8b200_subgrid_outer_flux_f32(int dim, int32_t nf, const int32_t* ranks, const int32_t* indices,
                                  const int32_t* face_neighbors, const float* face_normals, const float* face_surfaces,
                                  const int32_t* face_level_difference, const int32_t* face_neighbor_offset,
                                  const float* const* const* vars_all, float* const* const* flux_all, void* stream);
int t8b200_subgrid_outer_flux_f64(int dim, int32_t nf, const int32_t* ranks, const int32_t* indices,
                                  const int32_t* face_neighbors, const double* face_normals,
                                  const double* face_surfaces, const int32_t* face_level_difference,
                                  const int32_t* face_neighbor_offset, const double* const* const* vars_all,
                                  double* const* const* flux_all, void* stream);
/* Replaces compute_boundary_fluxes<<<B,(4,4)>>> (kernels.inl:913-1107, solver.inl:172): wall flux on the nb boundary
 * faces stored after the nf interior ones.  vars_own_tab / flux_own_tab: HOST arrays of 5 DEVICE one-entry tables
 * (this rank's pointer). */
int t8b200_subgrid_boundary_flux_f32(int dim, int32_t nf, int32_t nb, const int32_t* face_neighbors,
                                     const float* face_normals, const float* face_surfaces,
                                     const float* const* const* vars_own_tab, float* const* const* flux_own_tab,
                                     void* stream);
int t8b200_subgrid_boundary_flux_f64(int dim, int32_t nf, int32_t nb, const int32_t* face_neighbors,
                                     const double* face_normals, const double* face_surfaces,
                                     const double* const* const* vars_own_tab, double* const* const* flux_own_tab,
                                     void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * 2. B200-native fused path: connectivity re-laid out into per-chunk tiles ("plan"), one kernel per RK stage that
 *    stages a chunk of elements + halo in shared memory, evaluates every face of the chunk once, gathers the
 *    fluxes per element without atomics and applies the RK combination.  The flux accumulators never touch HBM.
 *    Replaces one stage of CompressibleEulerSolver::iterate (examples/compressible_euler/solver.cu:78-112).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct t8b200_plan t8b200_plan;

/* Builds the tile plan from the reference-layout connectivity (HOST pointers; same arrays as above, normals/areas
 * in the precision selected by is_f64).  ranks/indices may be NULL for a single rank without ghosts.
 * x_*: optional extra partition-boundary faces (ghost neighbour owned by a LOWER rank, which the reference assigns
 * to that rank, mesh_manager.inl:397); with them every rank evaluates all faces of its own elements
 * ("owner computes") and no remote atomics are needed.  n_xfaces may be 0. */
int t8b200_plan_create(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb,
                       const int32_t* face_neighbors, const void* face_normals, const void* face_surfaces,
                       const int32_t* ranks, const int32_t* indices, int32_t n_xfaces,
                       const int32_t* x_face_neighbors, const void* x_face_normals, const void* x_face_surfaces);
void t8b200_plan_destroy(t8b200_plan* plan);
/* info[0]=n_chunks, [1]=max halo per chunk, [2]=max faces per chunk, [3]=shared memory bytes per CTA,
 * [4]=device bytes held by the plan, [5]=face records (faces counted once per chunk they touch),
 * [6]=total halo entries, [7]=elements per chunk */
int t8b200_plan_info(const t8b200_plan* plan, int64_t info[8]);
/* The plan builder without a device (CPU-side checks of the host logic; the reference has no counterpart): same
 * arguments as t8b200_plan_create, nothing is uploaded, the plan cannot be launched.  t8b200_plan_host_array returns a
 * pointer into the plan's host arrays: which = 0 chunk headers (8 int32 per chunk: first element, count,
 * halo | faces << 16, x-end | y-end << 16, z-end, overflow offset index or -1, overflow entry index, uniform area index
 * or -1), 1 halo element indices (stride info[..] per chunk, -1 padded), 2 halo ranks, 3 face records (left slot bits
 * 0-13, axis bits 14-15, right slot or wall code bits 16-31), 4 area index per record, 5 element -> face table (8 uint16
 * per element: record << 1 | side, 0xFFFF none), 6 / 7 overflow offsets / entries, 8 area table, 9-12 normals and areas
 * of general plans (as double). */
int t8b200_plan_create_host(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb,
                            const int32_t* face_neighbors, const void* face_normals, const void* face_surfaces,
                            const int32_t* ranks, const int32_t* indices, int32_t n_xfaces,
                            const int32_t* x_face_neighbors, const void* x_face_normals, const void* x_face_surfaces);
/* Check access to the DEVICE builder's logic without a device: a host-only plan produced by the per-block program the
 * CUDA threads of t8b200_plan_create_device run (csrc/plan_block.cuh), driven by host loops.  Same arguments and arrays
 * as t8b200_plan_create_host; multi-rank plans stop at the chunk arrays (arrays 0-12). */
int t8b200_plan_create_block_program_host(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf,
                                          int32_t nb, const int32_t* nbr, const void* normals, const void* areas,
                                          const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                                          const void* xnormals, const void* xareas);
int t8b200_plan_host_array(const t8b200_plan* plan, int which, const void** data, int64_t* count, int* elem_bytes);

/* Ghost tail.  The reference keeps no ghost layer: kernels dereference the owner's arrays directly
 * (t8gpu/memory/shared_device_vector.h:25-29), which over NVLink exposes the peer-load latency inside every chunk that
 * touches the partition boundary (measured: +8 % per step on 2 GPUs).  A plan created with *_ghost_tail gives every ghost
 * element its chunks read a LOCAL copy at index n_local + j of this rank's own rows (rows need n_local +
 * t8b200_plan_ghost_tail_count() entries), the halo entries point there, and t8b200_ghost_pull_* fills the copies from
 * the peers' rows in one bandwidth-bound kernel before each stage: no pack / unpack on the owner's side, no ghost
 * buffers to exchange, the [var][rank] tables (get_all_variables) stay the interface.  The stage entry points then need no
 * in_all (it is ignored).  Arguments as t8b200_plan_create. */
int t8b200_plan_create_ghost_tail(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf,
                                  int32_t nb, const int32_t* face_neighbors, const void* face_normals,
                                  const void* face_surfaces, const int32_t* ranks, const int32_t* indices,
                                  int32_t n_xfaces, const int32_t* x_face_neighbors, const void* x_face_normals,
                                  const void* x_face_surfaces);
int t8b200_plan_create_ghost_tail_host(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf,
                                       int32_t nb, const int32_t* face_neighbors, const void* face_normals,
                                       const void* face_surfaces, const int32_t* ranks, const int32_t* indices,
                                       int32_t n_xfaces, const int32_t* x_face_neighbors, const void* x_face_normals,
                                       const void* x_face_surfaces);
/* tail entries of the plan (0 for plans without a ghost tail), -1 for a NULL plan; host arrays 17 / 18 of
 * t8b200_plan_host_array: owner rank / index in the owner's rows of every tail entry */
int64_t t8b200_plan_ghost_tail_count(const t8b200_plan* plan);
/* rows[k][n_local + j] = rows_all[k][rank_j][index_j] for every tail entry j and variable k < nvar (<= 8).
 * rows: HOST array of nvar DEVICE pointers (this rank's rows of one step); rows_all: HOST array of nvar DEVICE tables
 * (one pointer per rank; MemoryAccessorAll).  The owners' rows must be complete (t8b200_peer_barrier before). */
int t8b200_ghost_pull_f32(const t8b200_plan* plan, int nvar, float* const* rows, const float* const* const* rows_all,
                          void* stream);
int t8b200_ghost_pull_f64(const t8b200_plan* plan, int nvar, double* const* rows, const double* const* const* rows_all,
                          void* stream);

/* The plan built ON THE DEVICE from DEVICE connectivity arrays (same arguments as t8b200_plan_create, all pointers device
 * pointers): no device -> host copy of the connectivity, no host loop over the faces.
 *   - Meshes whose every block of 256 consecutive elements is an 8 x 8 x 4 box of same-size hexahedra with 256 single
 *     same-size face neighbours (uniform forests and brick partitions: the arrays t8b200_cartesian_*_connectivity leaves
 *     on the device): three kernels and a sort of the ghost keys.
 *   - Any other mesh (hanging faces, walls, general normals, ragged ends, blocks that must be split): one CUDA thread
 *     per block of 256 elements runs the block program of csrc/plan_block.cuh between data-parallel passes over the
 *     faces and the chunks; device -> host traffic is counters, the <= 256 distinct areas and two flag bytes per chunk.
 * Every array of the plan equals the host builder's, entry for entry (t8b200_plan_device_bytes, numbered as
 * t8b200_plan_host_array).  Returns cudaErrorNotSupported (801) only for n_local == 0.  The plan serves
 * t8b200_fused_stage_*, t8b200_ghost_pull_* and t8b200_gradient_criteria_*.  ghost_tail != 0: as
 * t8b200_plan_create_ghost_tail.
 * Replaces the upload half of MeshManager::compute_connectivity_information (t8gpu/mesh/mesh_manager.inl:442-480). */
int t8b200_plan_create_device(t8b200_plan** out, int is_f64, int ghost_tail, int64_t n_local, int64_t n_ghost,
                              int32_t nf, int32_t nb, const int32_t* face_neighbors, const void* face_normals,
                              const void* face_surfaces, const int32_t* ranks, const int32_t* indices,
                              int32_t n_xfaces, const int32_t* x_face_neighbors, const void* x_face_normals,
                              const void* x_face_surfaces, void* stream);
/* test access: one of the structured / ghost-tail DEVICE arrays of a plan copied to the host (which = 13, 14, 15, 17,
 * 18 as t8b200_plan_host_array numbers them); returns the element count (host_out may be NULL), -1 on error */
int64_t t8b200_plan_device_array(const t8b200_plan* plan, int which, int32_t* host_out, int64_t capacity);
/* test access to EVERY device array of a plan (which = 0 ... 19 as t8b200_plan_host_array), raw bytes in the device
 * element type; returns the byte count (host_out may be NULL), -1 on error */
int64_t t8b200_plan_device_bytes(const t8b200_plan* plan, int which, void* host_out, int64_t capacity_bytes);

/* The same exchange pushed by the owner: rows_all[k][dst_rank[e]][dst_idx[e]] = rows[k][src_idx[e]] for the n_send
 * entries of this rank's send list (DEVICE arrays; the pull lists of the peers -- host arrays 17 / 18 of their plans --
 * regrouped by owner and sorted by destination: what rank p pulls from this rank, with dst_idx = n_local(p) + j).  The
 * remote stores are consecutive (whole lines over NVLink, posted), only the local gather is scattered; a
 * t8b200_peer_barrier behind it publishes the copies.  rows: HOST array of nvar DEVICE pointers (this rank's rows of the
 * step that was just written); rows_all: HOST array of nvar DEVICE tables of the same step. */
int t8b200_ghost_push_f32(int nvar, int64_t n_send, const int32_t* src_idx, const int32_t* dst_rank,
                          const int32_t* dst_idx, const float* const* rows, float* const* const* rows_all, void* stream);
int t8b200_ghost_push_f64(int nvar, int64_t n_send, const int32_t* src_idx, const int32_t* dst_rank,
                          const int32_t* dst_idx, const double* const* rows, double* const* const* rows_all,
                          void* stream);
/* The push and the t8b200_peer_barrier that publishes it in ONE launch (the last CTA to finish its pushes runs the
 * barrier protocol): arguments of t8b200_ghost_push_* + those of t8b200_peer_barrier (value_dev != NULL: CFL slots and
 * maximum into out_max_dev, epoch from the CFL sequence; else stage slots / stage sequence).  counter_dev: DEVICE
 * unsigned, zero before the first call, left at zero by every call.  A rank with n_send == 0 still takes part. */
int t8b200_ghost_push_barrier_f32(int nvar, int64_t n_send, const int32_t* src_idx, const int32_t* dst_rank,
                                  const int32_t* dst_idx, const float* const* rows, float* const* const* rows_all,
                                  unsigned* counter_dev, int nranks, int rank, long long epoch, void* const* mailboxes_dev,
                                  const float* value_dev, float* out_max_dev, void* stream);
int t8b200_ghost_push_barrier_f64(int nvar, int64_t n_send, const int32_t* src_idx, const int32_t* dst_rank,
                                  const int32_t* dst_idx, const double* const* rows, double* const* const* rows_all,
                                  unsigned* counter_dev, int nranks, int rank, long long epoch, void* const* mailboxes_dev,
                                  const double* value_dev, double* out_max_dev, void* stream);

/* One fused RK stage.  in/prev/out: HOST arrays of 5 DEVICE pointers to this rank's arrays (stage input, U^n, stage
 * output).  in_all: HOST array of 5 DEVICE tables (one pointer per rank) for ghost reads, or NULL when the plan has
 * no ghosts.  vol: device, per element.  speed_max_dev: device scalar receiving max(|uHat|+aHat) over the faces of
 * this stage (zeroed by the call, then max-ed into by the kernel), or NULL. */
int t8b200_fused_stage_f32(const t8b200_plan* plan, int stage, const float* const* in,
                           const float* const* const* in_all, const float* const* prev, float* const* out,
                           const float* vol, float dt, float* speed_max_dev, void* stream);
int t8b200_fused_stage_f64(const t8b200_plan* plan, int stage, const double* const* in,
                           const double* const* const* in_all, const double* const* prev, double* const* out,
                           const double* vol, double dt, double* speed_max_dev, void* stream);

/* Fused Subgrid<4,4,4> (dim = 3) / Subgrid<4,4> (dim = 2) stage: replaces compute_inner_fluxes +
 * compute_boundary_fluxes + compute_outer_fluxes + subgrid::SSP_3RK_step{1,2,3} of one stage
 * (examples/subgrid/solver.inl:156-194) by ONE kernel without atomics: the subgrid mesh is handed to the tile-plan
 * stage kernel at cell level (256 consecutive cells per chunk).
 * The plan is built from the SubgridMeshConnectivityAccessor arrays (HOST pointers; normals with `dim` components per
 * face, offsets `dim` per face, t8gpu/mesh/subgrid_mesh_manager.h:108-126) and the per-element volumes (host, n_local;
 * they define the area of the faces between the cells of an element, kernels.inl:352-354, :542-544).
 * x_*: optional faces whose ghost neighbour is owned by a lower rank, as in t8b200_plan_create. */
typedef struct t8b200_subgrid_plan t8b200_subgrid_plan;
int  t8b200_subgrid_plan_create(t8b200_subgrid_plan** out, int is_f64, int dim, int64_t n_local, int64_t n_ghost,
                                int32_t nf, int32_t nb, const int32_t* face_neighbors, const void* face_normals,
                                const void* face_surfaces, const int32_t* face_level_difference,
                                const int32_t* face_neighbor_offset, const void* volumes, const int32_t* ranks,
                                const int32_t* indices, int32_t n_xfaces, const int32_t* x_face_neighbors,
                                const void* x_face_normals, const void* x_face_surfaces,
                                const int32_t* x_level_difference, const int32_t* x_neighbor_offset);
void t8b200_subgrid_plan_destroy(t8b200_subgrid_plan* plan);
/* host-only variant (see t8b200_plan_create_host) and access to the cell-level plan behind a subgrid plan */
int t8b200_subgrid_plan_create_host(t8b200_subgrid_plan** out, int is_f64, int dim, int64_t n_local, int64_t n_ghost,
                                    int32_t nf, int32_t nb, const int32_t* face_neighbors, const void* face_normals,
                                    const void* face_surfaces, const int32_t* face_level_difference,
                                    const int32_t* face_neighbor_offset, const void* volumes, const int32_t* ranks,
                                    const int32_t* indices, int32_t n_xfaces, const int32_t* x_face_neighbors,
                                    const void* x_face_normals, const void* x_face_surfaces,
                                    const int32_t* x_level_difference, const int32_t* x_neighbor_offset);
/* as t8b200_plan_create_block_program_host: the device builder's per-block program over the cell faces, on the host */
int t8b200_subgrid_plan_create_block_program_host(t8b200_subgrid_plan** out, int is_f64, int dim, int64_t n_local,
                                                  int64_t n_ghost, int32_t nf, int32_t nb, const int32_t* nbr,
                                                  const void* normals, const void* areas, const int32_t* level_diff,
                                                  const int32_t* offsets, const void* volumes, const int32_t* ranks,
                                                  const int32_t* indices, int32_t nx, const int32_t* xnbr,
                                                  const void* xnormals, const void* xareas, const int32_t* xld,
                                                  const int32_t* xoff);
/* ghost-tail variant (see t8b200_plan_create_ghost_tail); tail entries are CELLS: rows need n_local * 64 (16) +
 * t8b200_plan_ghost_tail_count(t8b200_subgrid_plan_base(plan)) entries, pulled with t8b200_ghost_pull_* on the base plan */
int t8b200_subgrid_plan_create_ghost_tail(t8b200_subgrid_plan** out, int is_f64, int dim, int64_t n_local,
                                          int64_t n_ghost, int32_t nf, int32_t nb, const int32_t* face_neighbors,
                                          const void* face_normals, const void* face_surfaces,
                                          const int32_t* face_level_difference, const int32_t* face_neighbor_offset,
                                          const void* volumes, const int32_t* ranks, const int32_t* indices,
                                          int32_t n_xfaces, const int32_t* x_face_neighbors, const void* x_face_normals,
                                          const void* x_face_surfaces, const int32_t* x_level_difference,
                                          const int32_t* x_neighbor_offset);
/* host-only variant of t8b200_subgrid_plan_create_ghost_tail (CPU-side checks, see t8b200_plan_create_host) */
int t8b200_subgrid_plan_create_ghost_tail_host(t8b200_subgrid_plan** out, int is_f64, int dim, int64_t n_local,
                                               int64_t n_ghost, int32_t nf, int32_t nb, const int32_t* nbr,
                                               const void* normals, const void* areas, const int32_t* level_diff,
                                               const int32_t* offsets, const void* volumes, const int32_t* ranks,
                                               const int32_t* indices, int32_t nx, const int32_t* xnbr,
                                               const void* xnormals, const void* xareas, const int32_t* xld,
                                               const int32_t* xoff);
/* Cell-level plan built on the device from DEVICE arrays (see t8b200_plan_create_device).  Subgrid<4,4,4> on forests
 * whose every group of 4 consecutive elements is a 2 x 2 x 1 block of same-size siblings surrounded by single same-level
 * elements (uniform forests and brick partitions): three kernels.  Any other forest (hanging faces, walls, Subgrid<4,4>):
 * the generic builder over the cell faces (16 / 4 per element face, 144 / 24 inside an element), one warp per 256
 * cells; every array equals t8b200_subgrid_plan_create's.  volumes: device, per element; offsets / x_offsets: dim per
 * face (same-level faces do not read them).  Arguments otherwise as t8b200_subgrid_plan_create. */
int t8b200_subgrid_plan_create_device(t8b200_subgrid_plan** out, int is_f64, int dim, int ghost_tail, int64_t n_local,
                                      int64_t n_ghost, int32_t nf, int32_t nb, const int32_t* face_neighbors,
                                      const void* face_normals, const void* face_surfaces,
                                      const int32_t* face_level_difference, const int32_t* face_neighbor_offset,
                                      const void* volumes, const int32_t* ranks, const int32_t* indices,
                                      int32_t n_xfaces, const int32_t* x_face_neighbors, const void* x_face_normals,
                                      const void* x_face_surfaces, const int32_t* x_level_difference,
                                      const int32_t* x_neighbor_offset, void* stream);
const t8b200_plan* t8b200_subgrid_plan_base(const t8b200_subgrid_plan* plan);
/* as t8b200_plan_info, counted in cells */
int t8b200_subgrid_plan_info(const t8b200_subgrid_plan* plan, int64_t info[8]);
/* in/prev/out: HOST arrays of 5 DEVICE pointers to this rank's cell arrays; in_all: tables for ghost reads or NULL;
 * vol: device, per ELEMENT (cell volume = vol/64 resp. vol/16, ssp_runge_kutta.inl:116). */
int t8b200_subgrid_fused_stage_f32(const t8b200_subgrid_plan* plan, int stage, const float* const* in,
                                   const float* const* const* in_all, const float* const* prev, float* const* out,
                                   const float* vol, float dt, void* stream);
int t8b200_subgrid_fused_stage_f64(const t8b200_subgrid_plan* plan, int stage, const double* const* in,
                                   const double* const* const* in_all, const double* const* prev, double* const* out,
                                   const double* vol, double dt, void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * 3. Device-side connectivity for uniform Cartesian periodic forests (quad / hex, one tree, Morton order):
 *    produces, bit for bit, the arrays MeshManager::compute_connectivity_information
 *    (t8gpu/mesh/mesh_manager.inl:332-481) uploads for t8_cmesh_new_periodic + t8_forest_new_uniform(level),
 *    for rank `rank` of `nranks` (contiguous SFC ranges, first element of rank p = floor(N p / P)).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t  n_local, n_ghost, n_faces, n_bfaces, n_xfaces;
  int32_t* ranks;          /* device, n_local + n_ghost */
  int32_t* indices;        /* device, n_local + n_ghost */
  int32_t* face_neighbors; /* device, 2 * n_faces */
  void*    face_normals;   /* device, 3 * n_faces, float or double */
  void*    face_surfaces;  /* device, n_faces */
  void*    volumes;        /* device, n_local */
  void*    centroids;      /* device, 3 * n_local (z = 0 in 2-D), already cast to float_type */
  int32_t* x_face_neighbors; /* device, 2 * n_xfaces: faces whose ghost neighbour is owned by a lower rank */
  void*    x_face_normals;
  void*    x_face_surfaces;
} t8b200_cart_conn;

int  t8b200_cartesian_uniform_connectivity(t8b200_cart_conn* out, int is_f64, int dim, int level, int nranks, int rank,
                                           void* stream);
/* Same for a periodic "brick" of bx*by*bz unit trees (tree id = x + bx (y + by z), elements ordered by tree, Morton
 * inside a tree): the weak-scaling meshes, one tree per GPU.  brick (2,2,2) at level L equals the level L+1 cube. */
int  t8b200_cartesian_brick_connectivity(t8b200_cart_conn* out, int is_f64, int dim, int level, int bx, int by, int bz,
                                         int nranks, int rank, void* stream);
void t8b200_cartesian_connectivity_free(t8b200_cart_conn* c);
/* The same arrays for an ADAPTIVE, 2:1 face-balanced one-tree quad / hex forest, from its leaves: keys_dev = Morton key
 * of every leaf's anchor at resolution 2^-20 per axis (bit b of coordinate d at position dim * b + d), levels_dev = its
 * level, both DEVICE arrays over all n_leaves leaves in SFC order (what t8_forest_get_element_in_tree enumerates;
 * 12 bytes per leaf instead of the face arrays).  Replaces the host loop of
 * MeshManager::compute_connectivity_information (t8gpu/mesh/mesh_manager.inl:358-440: t8_forest_leaf_face_neighbors per
 * face) bit for bit: hanging faces from the fine side (:411-424), ghost faces first and owned by the lower rank with
 * area / num_neighbors (:396-409), boundary faces behind the interior ones (:431-440), ghosts in SFC order; periodic = 0:
 * walls on the unit cube.  Rank `rank` of `nranks` holds the leaves [floor(N rank / P), floor(N (rank + 1) / P)).
 * Free with t8b200_cartesian_connectivity_free. */
int t8b200_forest_connectivity(t8b200_cart_conn* out, int is_f64, int dim, int periodic, int64_t n_leaves,
                               const uint64_t* keys_dev, const int32_t* levels_dev, int nranks, int rank, void* stream);
/* The same forest in the layout of SubgridMeshManager::compute_connectivity_information
 * (t8gpu/mesh/subgrid_mesh_manager.inl:788-905, add_face :560-786): `dim` normal components per face, and per interior /
 * x-face the level difference (<= 0: second element finer by that many levels) and the `dim` cell offsets of the face
 * inside the coarser element's 4^dim grid; the pair is swapped (normal flipped) when the neighbour is the finer one.
 * out->volumes are ELEMENT volumes.  Free `out` with t8b200_cartesian_connectivity_free and `info` with
 * t8b200_subgrid_face_info_free. */
typedef struct {
  int32_t* level_diff;   /* n_faces */
  int32_t* offsets;      /* n_faces * dim */
  int32_t* x_level_diff; /* n_xfaces */
  int32_t* x_offsets;    /* n_xfaces * dim */
} t8b200_subgrid_face_info;
int  t8b200_forest_subgrid_connectivity(t8b200_cart_conn* out, t8b200_subgrid_face_info* info, int is_f64, int dim,
                                        int periodic, int64_t n_leaves, const uint64_t* keys_dev,
                                        const int32_t* levels_dev, int nranks, int rank, void* stream);
void t8b200_subgrid_face_info_free(t8b200_subgrid_face_info* info);

/* -------------------------------------------------------------------------------------------------------------
 * 4. Cross-GPU sharing of the variable buffers, one process per GPU.  Replaces the MPI_Allgather of
 *    cudaIpcMemHandle_t in SharedDeviceVector (t8gpu/memory/shared_device_vector.inl:171-198): the 64-byte handles are
 *    exchanged by the caller (torch.distributed / NCCL here), the mapped pointers are peer pointers reached over
 *    NVLink.  All calls are synchronous.
 * ------------------------------------------------------------------------------------------------------------- */
int t8b200_shared_alloc(size_t bytes, void** dev_ptr, unsigned char handle[64]);
int t8b200_shared_open(const unsigned char handle[64], void** dev_ptr);
int t8b200_shared_close(void* dev_ptr);
int t8b200_shared_free(void* dev_ptr);

/* Stage barrier + max-reduction between the GPUs of one node over peer memory; stays on the stream.  Replaces the
 * cudaDeviceSynchronize() + MPI_Barrier pairs of iterate() (examples/compressible_euler/solver.cu:98-99, ...) and the
 * MPI_Allreduce(MAX) of compute_timestep (solver.cu:219-223).
 * mailboxes_dev: DEVICE array of nranks pointers; entry p is rank p's mailbox: 4 * nranks slots of 16 bytes,
 *   zero-initialised, allocated with t8b200_shared_alloc and mapped here with t8b200_shared_open.  Slots [0, 2 nranks)
 *   carry stage epochs, [2 nranks, 4 nranks) (value, epoch) pairs of the CFL reduction; each class is double-buffered by
 *   epoch parity (csrc/peer_sync.cuh).  Every rank stores (value, epoch) into its slot of every mailbox, then waits
 *   until all slots of its own mailbox carry `epoch`.
 * epoch: > 0, the same on all ranks, consecutive PER CLASS: calls without a value (and the stage kernels of
 *   t8b200_fused_stage_sync_*, which signal the same slots) share one sequence, calls with a value another.
 * value_dev: device scalar (float or double per value_is_f64) or NULL (pure barrier); out_max_dev: receives the
 *   maximum over the ranks, or NULL.  nranks <= 32. */
int t8b200_peer_barrier(int nranks, int rank, long long epoch, void* const* mailboxes_dev, const void* value_dev,
                        int value_is_f64, void* out_max_dev, void* stream);

/* CompressibleEulerSolver::compute_timestep's formula (examples/compressible_euler/solver.cu:225-228) on the device:
 *   *dt_dev = cfl * length / *speed_max_dev, capped by dt_cap when dt_cap > 0 (length = 0.5^max_level in the example).
 * With t8b200_fused_stage_sync_* reading dt from dt_dev, a time loop with the CFL rule needs no device -> host copy
 * and no host synchronisation per step (the reference's thrust::reduce returns through the host, solver.cu:214-217). */
int t8b200_timestep_f32(const float* speed_max_dev, float cfl, float length, float dt_cap, float* dt_dev, void* stream);
int t8b200_timestep_f64(const double* speed_max_dev, double cfl, double length, double dt_cap, double* dt_dev,
                        void* stream);

/* t8b200_fused_stage_* with (i) the time step read from device memory and (ii) the stage ordering between the GPUs done
 * by the stage kernel itself instead of a barrier between the launches:
 *   dt_dev != NULL: the kernels read dt from *dt_dev (written earlier on the stream, e.g. by t8b200_timestep_*); `dt` is
 *     ignored.
 *   sync != NULL (plans with ghosts): the chunks that read ghost elements ("partition-boundary chunks", scheduled first)
 *     wait until every peer has signalled `wait_epoch` (0: no wait), and the last of them to finish signals
 *     `signal_epoch` (0: no signal) to every peer; all other chunks neither wait nor signal, so they overlap the
 *     peers' skew.  Call pattern for consecutive stage launches k = 1, 2, 3, ... on every rank: wait_epoch = k - 1,
 *     signal_epoch = k (the very first launch after the state was written by other means must be preceded by a
 *     t8b200_peer_barrier, whose epoch continues the same sequence).  Replaces solver.cu:98-99,111-112,... as
 *     t8b200_peer_barrier does, minus one launch per stage and minus the lock step of the interior.
 *   No output may alias an input of the same call (in / prev / in_all rows vs out rows): cudaErrorInvalidValue. */
typedef struct {
  int          nranks, rank;
  void* const* mailboxes_dev; /* as t8b200_peer_barrier */
  unsigned*    counter_dev;   /* device, one zero-initialised unsigned owned by this rank */
} t8b200_stage_sync;
int t8b200_fused_stage_sync_f32(const t8b200_plan* plan, int stage, const float* const* in,
                                const float* const* const* in_all, const float* const* prev, float* const* out,
                                const float* vol, float dt, const float* dt_dev, float* speed_max_dev,
                                const t8b200_stage_sync* sync, long long wait_epoch, long long signal_epoch,
                                void* stream);
int t8b200_fused_stage_sync_f64(const t8b200_plan* plan, int stage, const double* const* in,
                                const double* const* const* in_all, const double* const* prev, double* const* out,
                                const double* vol, double dt, const double* dt_dev, double* speed_max_dev,
                                const t8b200_stage_sync* sync, long long wait_epoch, long long signal_epoch,
                                void* stream);
/* A stage of a ghost-tail plan in TWO launches, so that the exchange of the ghosts hides behind the interior:
 *   part 1  every chunk that reads no ghost copy (launched over all chunks; the partition-boundary ones leave at once).
 *           Needs nothing from the peers: the caller launches it right behind the previous stage.
 *   part 2  the partition-boundary chunks only; the caller launches it (typically on a second stream) after
 *           t8b200_peer_barrier + t8b200_ghost_pull_* of this stage.
 * Both parts together write exactly what t8b200_fused_stage_* writes (bitwise) and max into speed_max_dev, which the
 * caller zeroes before BOTH parts (they may run in either order on two streams).  For ghost-tail plans whose chunks are all structured (e.g. from t8b200_plan_create_device); other plans:
 * cudaErrorNotSupported, use t8b200_fused_stage_* after the pull.  dt_dev as in t8b200_fused_stage_sync_*. */
int t8b200_fused_stage_part_f32(const t8b200_plan* plan, int stage, int part, const float* const* in,
                                const float* const* prev, float* const* out, const float* vol, float dt,
                                const float* dt_dev, float* speed_max_dev, void* stream);
int t8b200_fused_stage_part_f64(const t8b200_plan* plan, int stage, int part, const double* const* in,
                                const double* const* prev, double* const* out, const double* vol, double dt,
                                const double* dt_dev, double* speed_max_dev, void* stream);
/* A stage of a ghost-tail plan with the PUSH folded into the stage kernel: the thread that writes an element other ranks
 * hold a ghost copy of also stores the new values into those copies (posted NVLink stores from the kernel's epilogue; no
 * push kernel, nothing to wait for inside the kernel).  out_all: HOST array of 5 DEVICE tables of the OUTPUT step;
 * send_off (n_local + 1) / send_rank / send_idx: DEVICE, CSR by element of the destinations (rank, index in that rank's
 * rows) -- the send list of t8b200_ghost_push_* grouped by source element.  A t8b200_peer_barrier behind the launch
 * publishes the copies.  For ghost-tail plans whose chunks are all structured; otherwise cudaErrorNotSupported. */
int t8b200_fused_stage_push_f32(const t8b200_plan* plan, int stage, const float* const* in, const float* const* prev,
                                float* const* out, float* const* const* out_all, const float* vol, float dt,
                                const float* dt_dev, float* speed_max_dev, const int32_t* send_off,
                                const int32_t* send_rank, const int32_t* send_idx, void* stream);
int t8b200_fused_stage_push_f64(const t8b200_plan* plan, int stage, const double* const* in, const double* const* prev,
                                double* const* out, double* const* const* out_all, const double* vol, double dt,
                                const double* dt_dev, double* speed_max_dev, const int32_t* send_off,
                                const int32_t* send_rank, const int32_t* send_idx, void* stream);
/* the same for the subgrid stage (no wave-speed reduction: the reference's subgrid solver has none) */
int t8b200_subgrid_fused_stage_sync_f32(const t8b200_subgrid_plan* plan, int stage, const float* const* in,
                                        const float* const* const* in_all, const float* const* prev,
                                        float* const* out, const float* vol, float dt, const float* dt_dev,
                                        const t8b200_stage_sync* sync, long long wait_epoch, long long signal_epoch,
                                        void* stream);
int t8b200_subgrid_fused_stage_sync_f64(const t8b200_subgrid_plan* plan, int stage, const double* const* in,
                                        const double* const* const* in_all, const double* const* prev,
                                        double* const* out, const double* vol, double dt, const double* dt_dev,
                                        const t8b200_stage_sync* sync, long long wait_epoch, long long signal_epoch,
                                        void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * 5. Device-side remap of variables and volumes after t8code adapt / partition (which stay on the host).
 *    Every value is written straight into the new arrays (no temporary + set_variable round trip).
 * ------------------------------------------------------------------------------------------------------------- */

/* Replaces adapt_variables_and_volume (t8gpu/mesh/mesh_manager.inl:164-193) when subgrid_dim = 0, and
 * adapt_variables + adapt_volume of the subgrid manager (t8gpu/mesh/subgrid_mesh_manager.inl:245-425) when
 * subgrid_dim = 3 (Subgrid<4,4,4>) or 2 (Subgrid<4,4>).
 * adapt_data: DEVICE, n_new + 1 entries, adapt_data[i] = first old element behind new element i
 *   (mesh_manager.inl:258-281).  Refined: every child copies its parent (subgrid: injection of the parent's octant),
 *   coarsened: mean of the family (subgrid: mean of 2^dim fine cells), else copy.  Volumes: x 1/8, 8 (elements and 3-D
 *   subgrids; the reference hard-codes the 3-D factors for MeshManager) or 1/4, 4 (2-D subgrids).
 * vars_old / vars_new: HOST arrays of nvar (<= 8) DEVICE pointers; vol_*: device, per element. */
int t8b200_adapt_remap_f32(int subgrid_dim, int nvar, int64_t n_new, const int32_t* adapt_data,
                           const float* const* vars_old, float* const* vars_new, const float* vol_old, float* vol_new,
                           void* stream);
int t8b200_adapt_remap_f64(int subgrid_dim, int nvar, int64_t n_new, const int32_t* adapt_data,
                           const double* const* vars_old, double* const* vars_new, const double* vol_old,
                           double* vol_new, void* stream);

/* Replaces partition_data (t8gpu/mesh/mesh_manager.inl:625-643; cells_per_element = 1) and
 * partition_variable_data + partition_volume_data (t8gpu/mesh/subgrid_mesh_manager.inl:1216-1283; cells_per_element
 * = 64 / 16): new element e <- old element indices[e] of rank ranks[e], read through the [var][rank] tables
 * (MemoryAccessorAll / get_all_volume; peer pointers over NVLink when the ranks are GPUs).
 * ranks / indices: DEVICE, n_new entries (what t8_forest_partition_data delivers); vars_new: HOST array of nvar DEVICE
 * pointers; vars_old_all: HOST array of nvar DEVICE tables; vol_new may be NULL (then vol_old_all is ignored). */
int t8b200_partition_remap_f32(int nvar, int64_t n_new, int cells_per_element, const int32_t* ranks,
                               const int32_t* indices, float* const* vars_new, const float* const* const* vars_old_all,
                               float* vol_new, const float* const* vol_old_all, void* stream);
int t8b200_partition_remap_f64(int nvar, int64_t n_new, int cells_per_element, const int32_t* ranks,
                               const int32_t* indices, double* const* vars_new,
                               const double* const* const* vars_old_all, double* vol_new,
                               const double* const* vol_old_all, void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * 6. Refinement indicators (the step before t8code adapt)
 * ------------------------------------------------------------------------------------------------------------- */

/* Replaces estimate_gradient<<<>>> + compute_refinement_criteria<<<>>> + the cudaMemset of the flux array
 * (examples/compressible_euler/kernels.cu:471-501, solver.cu:231-263):
 *   criteria[e] = sum over the interior faces of e of |rho_R - rho_L|, divided by cbrt(volume[e]).
 * Uses the tile plan (every rank sums all faces of its own elements: no remote atomics, fixed order, the flux array is
 * not touched).  rho: device, this rank's density; rho_all: device table of one pointer per rank (NULL without
 * ghosts); vol, criteria: device, per element. */
int t8b200_gradient_criteria_f32(const t8b200_plan* plan, const float* rho, const float* const* rho_all,
                                 const float* vol, float* criteria, void* stream);
int t8b200_gradient_criteria_f64(const t8b200_plan* plan, const double* rho, const double* const* rho_all,
                                 const double* vol, double* criteria, void* stream);

/* Replaces compute_refinement_criteria<SubgridType><<<>>> (examples/subgrid/kernels.inl:1109-1168, launched at
 * solver.inl:331-337): H1 seminorm of the density over the cells of each element / volume, summed in the reference's
 * loop order.  dim = 3: Subgrid<4,4,4>, dim = 2: Subgrid<4,4>.  rho: device, n_elements * 64 (16) cells. */
int t8b200_subgrid_criteria_f32(int dim, int64_t n_elements, const float* rho, const float* vol, float* criteria,
                                void* stream);
int t8b200_subgrid_criteria_f64(int dim, int64_t n_elements, const double* rho, const double* vol, double* criteria,
                                void* stream);

/* Cartesian Kelvin-Helmholtz initial state (examples/subgrid/solver.inl:36-56 / :82-103) sampled at n points
 * (device, 3 per point, float_type).  u: HOST array of 5 device pointers. */
int t8b200_init_kelvin_helmholtz_f32(int dim, int64_t n, const float* centers, float* const* u, void* stream);
int t8b200_init_kelvin_helmholtz_f64(int dim, int64_t n, const double* centers, double* const* u, void* stream);
/* Kelvin-Helmholtz initial state on the globe of the unstructured example (examples/compressible_euler/solver.cu:17-72,
 * a host lambda per element there) sampled at n element centroids (device, 3 per point, already cast to float_type as
 * at :32-34; r > 0 and not on the z axis).  Same expression types as the reference; agrees with the host evaluation to
 * the rounding of the device's sqrt / acos / asin / sin / cos / exp. */
int t8b200_init_spherical_kelvin_helmholtz_f32(int64_t n, const float* centers, float* const* u, void* stream);
int t8b200_init_spherical_kelvin_helmholtz_f64(int64_t n, const double* centers, double* const* u, void* stream);

/* -------------------------------------------------------------------------------------------------------------
 * 7. Output path of the subgrid manager (the step after the hot path when a run writes VTK)
 * ------------------------------------------------------------------------------------------------------------- */

/* Replaces column_major_to_z_order<<<N, Subgrid::block_size>>> + the device -> host copy of a float_type temporary +
 * the host loop widening float to double (t8gpu/mesh/subgrid_mesh_manager.inl:1007-1049, :1057-1064, :1106-1109):
 *   to[e * S + morton(i, j, k)] = (double) from[e * S + i + 4 j + 16 k],  S = 64 (dim 3) or 16 (dim 2),
 * morton = bit interleave x, y, z (x lowest) -- the leaf order of the forest refined twice more, which is the element
 * order t8_forest_write_vtk_ext expects.  from: device, n_elements * S cells of one variable; to: device, n_elements *
 * S doubles, must not alias from. */
int t8b200_subgrid_z_order_f32(int dim, int64_t n_elements, const float* from, double* to, void* stream);
int t8b200_subgrid_z_order_f64(int dim, int64_t n_elements, const double* from, double* to, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T8GPU_B200_H */
