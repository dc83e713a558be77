// Subgrid<4,4,4> / Subgrid<4,4> hot path for sm_100a.  C ABI in include/t8gpu_b200.h.
//
// Reference behaviour replaced (not translated):
//   examples/subgrid/kernels.inl:335-662   compute_inner_fluxes
//   examples/subgrid/kernels.inl:664-911   compute_outer_fluxes
//   examples/subgrid/kernels.inl:913-1107  compute_boundary_fluxes
//   examples/subgrid/solver.inl:152-266    iterate() schedule
// Cell (e,i,j,k) of a variable lives at  base[e*S + i + 4j + 16k]  (t8gpu/memory/subgrid_memory_manager.h:35-135).
#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

#include "../../include/t8gpu_b200.h"
#include "common.cuh"
#include "euler_flux.cuh"
#include "plan_emulate.cuh"
#include "subgrid_faces.cuh"
#include "tile_plan.cuh"

using namespace t8b200;

template <typename T>
struct SgOwn { T* p[5]; };
template <typename T>
struct SgOwnC { const T* p[5]; };
template <typename T>
struct SgAll { T* const* p[5]; };
template <typename T>
struct SgAllC { const T* const* p[5]; };

// ============================================================================================================
// 1. reference-shaped kernels
// ============================================================================================================

// 256 threads = 4 hexahedral (3-D) or 16 quadrilateral (2-D) elements, one thread per cell.  Each thread evaluates the
// faces towards its +x/+y/+z neighbour inside the element from cell quantities staged in shared memory, then every
// cell sums (-own +lower) and does ONE read-modify-write per variable on the global accumulator (the reference does 3
// per variable and recomputes the primitives of both cells for each of the three faces).
template <typename T, int DIM>
__global__ void __launch_bounds__(256) sg_inner_kernel(int64_t ne, const T* __restrict__ vol, SgOwnC<T> u, SgOwn<T> fl) {
  constexpr int S = DIM == 3 ? 64 : 16;
  constexpr int EPB = 256 / S;
  __shared__ T cq[NCELLQ][256];
  __shared__ T fx[DIM][5][256];
  const int     tid = threadIdx.x;
  const int64_t e   = (int64_t)blockIdx.x * EPB + tid / S;
  const int     c   = tid % S;
  const int     i = c & 3, j = (c >> 2) & 3, k = DIM == 3 ? c >> 4 : 0;
  const bool    on = e < ne;
  const int64_t g  = e * S + c;
  if (on) {
    Cell<T> q = to_cell(u.p[0][g], u.p[1][g], u.p[2][g], u.p[3][g], u.p[4][g]);
    cq[0][tid] = q.rho; cq[1][tid] = q.hx; cq[2][tid] = q.hy; cq[3][tid] = q.hz; cq[4][tid] = q.kp;
    cq[5][tid] = q.b;   cq[6][tid] = q.q;
  }
  __syncthreads();
  T surface = T(0);
  if (on) {
    T v = vol[e];
    if (DIM == 3) {
      T edge  = cbrt(v) / T(4);   // kernels.inl:352-354
      surface = edge * edge;
    } else {
      surface = sqrt(v) / T(4);   // kernels.inl:542-544
    }
  }
  const int ijk[3] = {i, j, k};
  const int st[3]  = {1, 4, 16};
#pragma unroll
  for (int ax = 0; ax < DIM; ax++) {
    if (on && ijk[ax] < 3) {
      Cell<T> L, R;
      int a = tid, b = tid + st[ax];
      L.rho = cq[0][a]; L.hx = cq[1][a]; L.hy = cq[2][a]; L.hz = cq[3][a]; L.kp = cq[4][a]; L.b = cq[5][a]; L.q = cq[6][a];
      R.rho = cq[0][b]; R.hx = cq[1][b]; R.hy = cq[2][b]; R.hz = cq[3][b]; R.kp = cq[4][b]; R.b = cq[5][b]; R.q = cq[6][b];
      T F[5];
      if (ax == 0) kepes_flux_n<T, 0>(L, R, T(0), T(0), T(0), F);
      else if (ax == 1) kepes_flux_n<T, 1>(L, R, T(0), T(0), T(0), F);
      else kepes_flux_n<T, 2>(L, R, T(0), T(0), T(0), F);
#pragma unroll
      for (int v = 0; v < 5; v++) fx[ax][v][tid] = F[v] * surface;
    }
  }
  __syncthreads();
  if (on) {
#pragma unroll
    for (int v = 0; v < 5; v++) {
      T acc = T(0);
#pragma unroll
      for (int ax = 0; ax < DIM; ax++) {
        if (ijk[ax] < 3) acc -= fx[ax][v][tid];
        if (ijk[ax] > 0) acc += fx[ax][v][tid - st[ax]];
      }
      fl.p[v][g] += acc;
    }
  }
}

// cell indices of face-thread (i,j) on both sides (kernels.inl:710-758)
template <int DIM>
__device__ __forceinline__ void sg_face_cells(const float n[3], const int off[3], int dstride, int i, int j, int& lc,
                                              int& rc) {
  int al[3] = {0, 0, 0}, si[3] = {0, 0, 0}, sj[3] = {0, 0, 0};
  if (n[0] == 1.0f) { al[0] = 3; si[1] = 1; sj[2] = 1; }
  if (n[0] == -1.0f) { si[1] = 1; sj[2] = 1; }
  if (n[1] == 1.0f) { al[1] = 3; si[0] = 1; sj[2] = 1; }
  if (n[1] == -1.0f) { si[0] = 1; sj[2] = 1; }
  if (DIM == 3) {
    if (n[2] == 1.0f) { al[2] = 3; si[0] = 1; sj[1] = 1; }
    if (n[2] == -1.0f) { si[0] = 1; sj[1] = 1; }
  }
  int l[3], r[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    l[d] = al[d] + i * si[d] + j * sj[d];
    r[d] = off[d] + dstride * (i * si[d] + j * sj[d]) / 2;
  }
  lc = l[0] + 4 * l[1] + 16 * l[2];
  rc = r[0] + 4 * r[1] + 16 * r[2];
}

// One thread per (face, i, j): 16 (3-D) or 4 (2-D) consecutive threads per face, 256-thread CTAs (the reference
// launches one 16- or 4-thread block per face).  Boundary faces (f >= nf) mirror the left state.
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
sg_outer_kernel(int nf, int nb, const int32_t* __restrict__ ranks, const int32_t* __restrict__ indices,
                const int32_t* __restrict__ nbr, const T* __restrict__ normals, const T* __restrict__ areas,
                const int32_t* __restrict__ level_diff, const int32_t* __restrict__ offsets, SgAllC<T> u, SgAll<T> fl,
                int first_face) {
  constexpr int S = DIM == 3 ? 64 : 16;
  constexpr int TPF = DIM == 3 ? 16 : 4;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int     f = first_face + (int)(t / TPF);
  if (f >= nf + nb) return;
  const int  q = (int)(t % TPF), i = q & 3, j = q >> 2;
  const bool boundary = f >= nf;
  int l = boundary ? nbr[2 * nf + (f - nf)] : nbr[2 * f];
  int r = boundary ? l : nbr[2 * f + 1];
  int lrk = 0, li = l, rrk = 0, ri = r;
  if (ranks) { lrk = ranks[l]; li = indices[l]; rrk = ranks[r]; ri = indices[r]; }
  T nx = normals[DIM * f], ny = normals[DIM * f + 1], nz = DIM == 3 ? normals[DIM * f + 2] : T(0);
  float nf3[3] = {(float)nx, (float)ny, (float)nz};
  int   off[3] = {0, 0, 0};
  int   dstride = 2;
  if (!boundary) {
    off[0] = offsets[DIM * f]; off[1] = offsets[DIM * f + 1]; off[2] = DIM == 3 ? offsets[DIM * f + 2] : 0;
    dstride = level_diff[f] == 0 ? 2 : 1;
  }
  int lc, rc;
  sg_face_cells<DIM>(nf3, off, dstride, i, j, lc, rc);
  const int64_t gl = (int64_t)li * S + lc, gr = (int64_t)ri * S + rc;
  Cell<T> L = to_cell(u.p[0][lrk][gl], u.p[1][lrk][gl], u.p[2][lrk][gl], u.p[3][lrk][gl], u.p[4][lrk][gl]);
  Cell<T> R = boundary ? mirror(L, nx, ny, nz)
                       : to_cell(u.p[0][rrk][gr], u.p[1][rrk][gr], u.p[2][rrk][gr], u.p[3][rrk][gr], u.p[4][rrk][gr]);
  T F[5];
  kepes_flux(L, R, nx, ny, nz, F);
  T surface = areas[f] / T(DIM == 3 ? 16 : 4);   // kernels.inl:786-787, :894
#pragma unroll
  for (int v = 0; v < 5; v++) {
    T w = F[v] * surface;
    atomicAdd(&fl.p[v][lrk][gl], -w);
    if (!boundary) atomicAdd(&fl.p[v][rrk][gr], w);
  }
}

template <typename T>
static int sg_inner_impl(int dim, int64_t ne, const T* vol, const T* const* vars, T* const* flux, void* stream) {
  if ((dim != 2 && dim != 3) || ne < 0 || !vars || !flux || (ne > 0 && !vol)) return cudaErrorInvalidValue;
  if (ne == 0) return 0;
  SgOwnC<T> u;
  SgOwn<T>  f;
  for (int k = 0; k < 5; k++) { u.p[k] = vars[k]; f.p[k] = flux[k]; }
  const int epb = dim == 3 ? 4 : 16;
  unsigned  blocks = (unsigned)((ne + epb - 1) / epb);
  if (dim == 3) sg_inner_kernel<T, 3><<<blocks, 256, 0, (cudaStream_t)stream>>>(ne, vol, u, f);
  else sg_inner_kernel<T, 2><<<blocks, 256, 0, (cudaStream_t)stream>>>(ne, vol, u, f);
  return cudaGetLastError();
}

template <typename T>
static int sg_outer_impl(int dim, int32_t nf, int32_t nb, int32_t first, int32_t count, const int32_t* ranks,
                         const int32_t* indices, const int32_t* nbr, const T* normals, const T* areas,
                         const int32_t* level_diff, const int32_t* offsets, const T* const* const* vars_all,
                         T* const* const* flux_all, void* stream) {
  if ((dim != 2 && dim != 3) || nf < 0 || nb < 0 || first < 0 || count < 0 || first + count > nf + nb || !vars_all ||
      !flux_all)
    return cudaErrorInvalidValue;
  if (count == 0) return 0;
  if (!nbr || !normals || !areas || (first < nf && (!level_diff || !offsets))) return cudaErrorInvalidValue;
  SgAllC<T> u;
  SgAll<T>  f;
  for (int k = 0; k < 5; k++) { u.p[k] = vars_all[k]; f.p[k] = flux_all[k]; }
  const int tpf = dim == 3 ? 16 : 4;
  unsigned  blocks = (unsigned)(((int64_t)count * tpf + 255) / 256);
  // the kernel stops at nf + nb_eff: 0 boundary faces when only interior faces [0, nf) are requested
  const int nb_eff = first + count <= nf ? 0 : first + count - nf;
  if (dim == 3)
    sg_outer_kernel<T, 3><<<blocks, 256, 0, (cudaStream_t)stream>>>(nf, nb_eff, ranks, indices, nbr, normals, areas,
                                                                   level_diff, offsets, u, f, first);
  else
    sg_outer_kernel<T, 2><<<blocks, 256, 0, (cudaStream_t)stream>>>(nf, nb_eff, ranks, indices, nbr, normals, areas,
                                                                   level_diff, offsets, u, f, first);
  return cudaGetLastError();
}

extern "C" {

int t8b200_subgrid_inner_flux_f32(int dim, int64_t ne, const float* vol, const float* const* vars, float* const* flux,
                                  void* stream) {
  return sg_inner_impl<float>(dim, ne, vol, vars, flux, stream);
}
int t8b200_subgrid_inner_flux_f64(int dim, int64_t ne, const double* vol, const double* const* vars,
                                  double* const* flux, void* stream) {
  return sg_inner_impl<double>(dim, ne, vol, vars, flux, stream);
}
int t8b200_subgrid_outer_flux_f32(int dim, int32_t nf, const int32_t* ranks, const int32_t* indices,
                                  const int32_t* nbr, const float* normals, const float* areas,
                                  const int32_t* level_diff, const int32_t* offsets, const float* const* const* vars_all,
                                  float* const* const* flux_all, void* stream) {
  return sg_outer_impl<float>(dim, nf, 0, 0, nf, ranks, indices, nbr, normals, areas, level_diff, offsets, vars_all,
                              flux_all, stream);
}
int t8b200_subgrid_outer_flux_f64(int dim, int32_t nf, const int32_t* ranks, const int32_t* indices,
                                  const int32_t* nbr, const double* normals, const double* areas,
                                  const int32_t* level_diff, const int32_t* offsets,
                                  const double* const* const* vars_all, double* const* const* flux_all, void* stream) {
  return sg_outer_impl<double>(dim, nf, 0, 0, nf, ranks, indices, nbr, normals, areas, level_diff, offsets, vars_all,
                               flux_all, stream);
}
// boundary faces: own arrays only; tables of one rank are built on the fly from the own pointers
int t8b200_subgrid_boundary_flux_f32(int dim, int32_t nf, int32_t nb, const int32_t* nbr, const float* normals,
                                     const float* areas, const float* const* const* vars_own_tab,
                                     float* const* const* flux_own_tab, void* stream) {
  return sg_outer_impl<float>(dim, nf, nb, nf, nb, nullptr, nullptr, nbr, normals, areas, nullptr, nullptr, vars_own_tab,
                              flux_own_tab, stream);
}
int t8b200_subgrid_boundary_flux_f64(int dim, int32_t nf, int32_t nb, const int32_t* nbr, const double* normals,
                                     const double* areas, const double* const* const* vars_own_tab,
                                     double* const* const* flux_own_tab, void* stream) {
  return sg_outer_impl<double>(dim, nf, nb, nf, nb, nullptr, nullptr, nbr, normals, areas, nullptr, nullptr,
                               vars_own_tab, flux_own_tab, stream);
}
}

// ============================================================================================================
// 2. fused Subgrid<4,4,4> / Subgrid<4,4> stage: owner-computes, no atomics, fluxes never reach HBM
// ============================================================================================================
//
// The subgrid mesh is handed to the tile-plan stage kernel (fused.cu) at CELL level: every cell is an "element" of the
// plan, a chunk of 256 consecutive cells is 4 hexahedral (16 quadrilateral) elements -- for same-level siblings an
// 8x8x4 (16x16) box of cells, the same tile shape as on an unstructured Morton-ordered hex (quad) forest.  The face
// source below enumerates
//   * the faces between the cells of one element (compute_inner_fluxes, kernels.inl:335-662): normal +e_axis, area
//     (cbrt(vol)/4)^2 in 3-D (kernels.inl:352-354), sqrt(vol)/4 in 2-D (:542-544);
//   * 16 (4) cell faces per element face (compute_outer_fluxes, kernels.inl:664-911): left cell from the face normal,
//     right cell r = anchor + (i * stride) / 2 (kernels.inl:756-758, stride 2 for equal levels, 1 towards a coarser
//     neighbour: a coarse cell receives its four fine sub-faces as four faces), area face_surface/16 (/4);
//   * 16 (4) wall faces per boundary face (compute_boundary_fluxes, kernels.inl:913-1107).
// Both owners of a face evaluate it in the same canonical orientation from the same two cells, so the scheme stays
// exactly conservative without atomics and is deterministic.  Cell volume = vol[element] / 64 (/16)
// (ssp_runge_kutta.inl:116).  Replaces inner + boundary + outer + SSP_3RK_step (examples/subgrid/solver.inl:156-194).

struct t8b200_subgrid_plan {
  t8b200_plan* plan = nullptr;
  int          dim  = 3;
};

t8b200_subgrid_plan* t8b_wrap_subgrid_plan(t8b200_plan* P, int dim) {
  auto* SP = new t8b200_subgrid_plan();
  SP->plan = P;
  SP->dim  = dim;
  return SP;
}

template <typename T, int DIM>
static int sg_plan_build_dim(t8b200_subgrid_plan* SP, int host_only, int64_t n_local, int64_t n_ghost, int32_t nf,
                             int32_t nb, const int32_t* nbr, const T* normals, const T* areas,
                             const int32_t* level_diff, const int32_t* offsets, const T* vol, const int32_t* ranks,
                             const int32_t* indices, int32_t nx, const int32_t* xnbr, const T* xnormals,
                             const T* xareas, const int32_t* xld, const int32_t* xoff);

template <typename T>
static int sg_plan_build(t8b200_subgrid_plan* SP, int host_only, int dim, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb,
                         const int32_t* nbr, const T* normals, const T* areas, const int32_t* level_diff,
                         const int32_t* offsets, const T* vol, const int32_t* ranks, const int32_t* indices, int32_t nx,
                         const int32_t* xnbr, const T* xnormals, const T* xareas, const int32_t* xld,
                         const int32_t* xoff) {
  if (dim == 2) return sg_plan_build_dim<T, 2>(SP, host_only, n_local, n_ghost, nf, nb, nbr, normals, areas, level_diff,
                                               offsets, vol, ranks, indices, nx, xnbr, xnormals, xareas, xld, xoff);
  return sg_plan_build_dim<T, 3>(SP, host_only, n_local, n_ghost, nf, nb, nbr, normals, areas, level_diff, offsets, vol,
                                 ranks, indices, nx, xnbr, xnormals, xareas, xld, xoff);
}

template <typename T, int DIM>
static int sg_plan_build_dim(t8b200_subgrid_plan* SP, int host_only, int64_t n_local, int64_t n_ghost, int32_t nf,
                             int32_t nb, const int32_t* nbr, const T* normals, const T* areas,
                             const int32_t* level_diff, const int32_t* offsets, const T* vol, const int32_t* ranks,
                             const int32_t* indices, int32_t nx, const int32_t* xnbr, const T* xnormals,
                             const T* xareas, const int32_t* xld, const int32_t* xoff) {
  constexpr int dim = DIM;
  // area of the faces between the cells, once per element (the plan asks for the geometry of a face several times)
  std::vector<T> inner((size_t)n_local);
  parallel_ranges(n_local, plan_threads(), [&](int, int64_t e0, int64_t e1) {
    for (int64_t e = e0; e < e1; e++) inner[e] = subgrid_inner_area<T, DIM>(vol[e]);
  });
  SubgridFaces<T, DIM> src{n_local, nf, nb, nx, nbr, normals, areas, level_diff, offsets, vol, ranks, indices, xnbr,
                           xnormals, xareas, xld, xoff, inner.data()};
  if ((n_local + n_ghost) * src.S() > 0x7FFFFFF0LL) return cudaErrorInvalidValue;
  t8b200_plan* P = new t8b200_plan();
  SP->plan       = P;
  P->is_f64      = sizeof(T) == 8;
  P->host_only   = host_only & 1;   // flags: bit 0 host-only plan, bit 1 ghost tail, bit 2 block program on the host
  P->ghost_tail  = (host_only >> 1) & 1;
  P->vol_shift   = dim == 3 ? 6 : 4;
  P->vol_scale   = dim == 3 ? 1.0 / 64.0 : 1.0 / 16.0;
  if ((host_only & 5) == 5) return plan_build_emulated<T>(P, n_local * src.S(), n_ghost > 0, src);
  return plan_build<T>(P, n_local * src.S(), n_ghost > 0, src);
}

extern "C" {

static int sg_plan_create_impl(t8b200_subgrid_plan** out, int host_only, int is_f64, int dim, int64_t n_local, int64_t n_ghost,
                               int32_t nf, int32_t nb, const int32_t* nbr, const void* normals, const void* areas,
                               const int32_t* level_diff, const int32_t* offsets, const void* volumes,
                               const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                               const void* xnormals, const void* xareas, const int32_t* xld, const int32_t* xoff) {
  if (!out || (dim != 2 && dim != 3) || n_local < 0 || n_ghost < 0 || nf < 0 || nb < 0 || nx < 0)
    return cudaErrorInvalidValue;
  if (nf + nb > 0 && (!nbr || !normals || !areas)) return cudaErrorInvalidValue;
  if (nf > 0 && (!level_diff || !offsets)) return cudaErrorInvalidValue;
  if (nx > 0 && (!xnbr || !xnormals || !xareas || !xld || !xoff)) return cudaErrorInvalidValue;
  if (n_local > 0 && !volumes) return cudaErrorInvalidValue;
  if (n_ghost > 0 && (!ranks || !indices)) return cudaErrorInvalidValue;
  auto* SP = new t8b200_subgrid_plan();
  SP->dim  = dim;
  int rc = is_f64 ? sg_plan_build<double>(SP, host_only, dim, n_local, n_ghost, nf, nb, nbr, (const double*)normals,
                                          (const double*)areas, level_diff, offsets, (const double*)volumes, ranks,
                                          indices, nx, xnbr, (const double*)xnormals, (const double*)xareas, xld, xoff)
                  : sg_plan_build<float>(SP, host_only, dim, n_local, n_ghost, nf, nb, nbr, (const float*)normals,
                                         (const float*)areas, level_diff, offsets, (const float*)volumes, ranks,
                                         indices, nx, xnbr, (const float*)xnormals, (const float*)xareas, xld, xoff);
  if (rc) {
    t8b200_subgrid_plan_destroy(SP);
    return rc;
  }
  *out = SP;
  return 0;
}
int t8b200_subgrid_plan_create(t8b200_subgrid_plan** out, int is_f64, int dim, int64_t n_local, int64_t n_ghost,
                               int32_t nf, int32_t nb, const int32_t* nbr, const void* normals, const void* areas,
                               const int32_t* level_diff, const int32_t* offsets, const void* volumes,
                               const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                               const void* xnormals, const void* xareas, const int32_t* xld, const int32_t* xoff) {
  return sg_plan_create_impl(out, 0, is_f64, dim, n_local, n_ghost, nf, nb, nbr, normals, areas, level_diff, offsets,
                             volumes, ranks, indices, nx, xnbr, xnormals, xareas, xld, xoff);
}
int t8b200_subgrid_plan_create_host(t8b200_subgrid_plan** out, int is_f64, int dim, int64_t n_local, int64_t n_ghost,
                                    int32_t nf, int32_t nb, const int32_t* nbr, const void* normals,
                                    const void* areas, const int32_t* level_diff, const int32_t* offsets,
                                    const void* volumes, const int32_t* ranks, const int32_t* indices, int32_t nx,
                                    const int32_t* xnbr, const void* xnormals, const void* xareas, const int32_t* xld,
                                    const int32_t* xoff) {
  return sg_plan_create_impl(out, 1, is_f64, dim, n_local, n_ghost, nf, nb, nbr, normals, areas, level_diff, offsets,
                             volumes, ranks, indices, nx, xnbr, xnormals, xareas, xld, xoff);
}
int t8b200_subgrid_plan_create_block_program_host(t8b200_subgrid_plan** out, int is_f64, int dim, int64_t n_local,
                                                  int64_t n_ghost, int32_t nf, int32_t nb, const int32_t* nbr,
                                                  const void* normals, const void* areas, const int32_t* level_diff,
                                                  const int32_t* offsets, const void* volumes, const int32_t* ranks,
                                                  const int32_t* indices, int32_t nx, const int32_t* xnbr,
                                                  const void* xnormals, const void* xareas, const int32_t* xld,
                                                  const int32_t* xoff) {
  return sg_plan_create_impl(out, 5, is_f64, dim, n_local, n_ghost, nf, nb, nbr, normals, areas, level_diff, offsets,
                             volumes, ranks, indices, nx, xnbr, xnormals, xareas, xld, xoff);
}
int t8b200_subgrid_plan_create_ghost_tail(t8b200_subgrid_plan** out, int is_f64, int dim, int64_t n_local,
                                          int64_t n_ghost, int32_t nf, int32_t nb, const int32_t* nbr,
                                          const void* normals, const void* areas, const int32_t* level_diff,
                                          const int32_t* offsets, const void* volumes, const int32_t* ranks,
                                          const int32_t* indices, int32_t nx, const int32_t* xnbr, const void* xnormals,
                                          const void* xareas, const int32_t* xld, const int32_t* xoff) {
  return sg_plan_create_impl(out, 2, is_f64, dim, n_local, n_ghost, nf, nb, nbr, normals, areas, level_diff, offsets,
                             volumes, ranks, indices, nx, xnbr, xnormals, xareas, xld, xoff);
}
int t8b200_subgrid_plan_create_ghost_tail_host(t8b200_subgrid_plan** out, int is_f64, int dim, int64_t n_local,
                                               int64_t n_ghost, int32_t nf, int32_t nb, const int32_t* nbr,
                                               const void* normals, const void* areas, const int32_t* level_diff,
                                               const int32_t* offsets, const void* volumes, const int32_t* ranks,
                                               const int32_t* indices, int32_t nx, const int32_t* xnbr,
                                               const void* xnormals, const void* xareas, const int32_t* xld,
                                               const int32_t* xoff) {
  return sg_plan_create_impl(out, 3, is_f64, dim, n_local, n_ghost, nf, nb, nbr, normals, areas, level_diff, offsets,
                             volumes, ranks, indices, nx, xnbr, xnormals, xareas, xld, xoff);
}
const t8b200_plan* t8b200_subgrid_plan_base(const t8b200_subgrid_plan* SP) { return SP ? SP->plan : nullptr; }
void t8b200_subgrid_plan_destroy(t8b200_subgrid_plan* SP) {
  if (!SP) return;
  t8b_plan_free(SP->plan);
  delete SP;
}
int t8b200_subgrid_plan_info(const t8b200_subgrid_plan* SP, int64_t info[8]) {
  if (!SP) return cudaErrorInvalidValue;
  return t8b200_plan_info(SP->plan, info);
}
int t8b200_subgrid_fused_stage_f32(const t8b200_subgrid_plan* SP, int stage, const float* const* in,
                                   const float* const* const* in_all, const float* const* prev, float* const* out,
                                   const float* vol, float dt, void* stream) {
  if (!SP) return cudaErrorInvalidValue;
  return t8b_fused_stage_run<float>(SP->plan, stage, in, in_all, prev, out, vol, dt, nullptr, stream);
}
int t8b200_subgrid_fused_stage_f64(const t8b200_subgrid_plan* SP, int stage, const double* const* in,
                                   const double* const* const* in_all, const double* const* prev, double* const* out,
                                   const double* vol, double dt, void* stream) {
  if (!SP) return cudaErrorInvalidValue;
  return t8b_fused_stage_run<double>(SP->plan, stage, in, in_all, prev, out, vol, dt, nullptr, stream);
}
int t8b200_subgrid_fused_stage_sync_f32(const t8b200_subgrid_plan* SP, int stage, const float* const* in,
                                        const float* const* const* in_all, const float* const* prev,
                                        float* const* out, const float* vol, float dt, const float* dt_dev,
                                        const t8b200_stage_sync* sync, long long wait_epoch, long long signal_epoch,
                                        void* stream) {
  if (!SP) return cudaErrorInvalidValue;
  return t8b_fused_stage_run<float>(SP->plan, stage, in, in_all, prev, out, vol, dt, nullptr, stream, dt_dev, sync,
                                    wait_epoch, signal_epoch);
}
int t8b200_subgrid_fused_stage_sync_f64(const t8b200_subgrid_plan* SP, int stage, const double* const* in,
                                        const double* const* const* in_all, const double* const* prev,
                                        double* const* out, const double* vol, double dt, const double* dt_dev,
                                        const t8b200_stage_sync* sync, long long wait_epoch, long long signal_epoch,
                                        void* stream) {
  if (!SP) return cudaErrorInvalidValue;
  return t8b_fused_stage_run<double>(SP->plan, stage, in, in_all, prev, out, vol, dt, nullptr, stream, dt_dev, sync,
                                     wait_epoch, signal_epoch);
}
}
