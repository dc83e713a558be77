// Subgrid<4,4,4> / Subgrid<4,4> hot path for sm_100a.  C ABI in include/t8gpu_b200.h.
//
// Reference behaviour replaced (not translated):
//   examples/subgrid/kernels.inl:335-662   compute_inner_fluxes
//   examples/subgrid/kernels.inl:664-911   compute_outer_fluxes
//   examples/subgrid/kernels.inl:913-1107  compute_boundary_fluxes
//   examples/subgrid/solver.inl:152-266    iterate() schedule
// Cell (e,i,j,k) of a variable lives at  base[e*S + i + 4j + 16k]  (t8gpu/memory/subgrid_memory_manager.h:35-135).
#include <algorithm>
#include <vector>

#include "../../include/t8gpu_b200.h"
#include "common.cuh"
#include "euler_flux.cuh"

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
      kepes_flux(L, R, ax == 0 ? T(1) : T(0), ax == 1 ? T(1) : T(0), ax == 2 ? T(1) : T(0), F);
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
// 2. fused Subgrid<4,4,4> stage: owner-computes, no atomics, fluxes never reach HBM
// ============================================================================================================
//
// CTA = 4 consecutive elements (256 cells, one thread per cell).  Each element's cells plus a one-cell halo layer on
// its 6 faces are staged as a padded 6x6x6 block of per-cell quantities in shared memory.  The halo of a face comes
// from the same-level neighbour's boundary layer, from the coarser neighbour's cells by injection (exactly the
// reference's r = anchor + (i*stride)/2 rule, kernels.inl:756-758), or from the mirrored own cell at a wall.  Every
// face of the 5 planes per axis is then evaluated ONCE per element with the canonical orientation (lower cell = left,
// normal = +axis): the two elements sharing a face evaluate bit-identical fluxes from the same two cells, so the
// scheme stays exactly conservative without atomics and is deterministic.  Where the neighbour is FINER (2:1 hanging
// face seen from the coarse side) the face flux of a coarse boundary cell is the sum of its four fine sub-faces.
// Replaces inner + boundary + outer + SSP_3RK_step (examples/subgrid/solver.inl:156-194).

struct t8b200_subgrid_plan {
  int      is_f64 = 0, multi = 0;
  int64_t  ne = 0, dev_bytes = 0;
  uint8_t* kind = nullptr;   // ne*6 : 0 wall, 1 same level, 2 coarser, 3 finer
  uint8_t* quad = nullptr;   // ne*6 : kind 2: quarter of the coarse face (a | b << 1)
  int32_t* nid  = nullptr;   // ne*6 : neighbour element (index in its owner's arrays) or row of fine_id
  int32_t* nrk  = nullptr;   // ne*6 : owner rank (multi)
  void*    aout = nullptr;   // ne*6 : cell-face area of that side ( = face_surface / 16 )
  int32_t* fine_id = nullptr;  // 4 per hanging coarse face
  int32_t* fine_rk = nullptr;
};

template <typename T>
struct SgFusedArgs {
  const uint8_t* kind;
  const uint8_t* quad;
  const int32_t* nid;
  const int32_t* nrk;
  const T*       aout;
  const int32_t* fine_id;
  const int32_t* fine_rk;
  const T*        in[5];
  const T* const* in_all[5];
  const T*        prev[5];
  T*              out[5];
  const T*        vol;
  T               dt;
  int64_t         ne;
  int             stage, multi;
};

namespace {
constexpr int G   = 4;      // elements per CTA
constexpr int PS  = 216;    // padded 6^3 slots per element
constexpr int NFA = 80;     // faces per axis per element (5 planes x 16)

__device__ __forceinline__ int pslot(int i, int j, int k) { return i + 6 * j + 36 * k; }  // padded coords 0..5

// (axis, plane-normal coordinate x, tangential a, b) -> padded coords
__device__ __forceinline__ int pslot_ax(int ax, int x, int a, int b) {
  return ax == 0 ? pslot(x, a + 1, b + 1) : (ax == 1 ? pslot(a + 1, x, b + 1) : pslot(a + 1, b + 1, x));
}
// cell index inside an element from (axis, x along axis, tangential a, b), unpadded 0..3
__device__ __forceinline__ int cell_ax(int ax, int x, int a, int b) {
  return ax == 0 ? x + 4 * a + 16 * b : (ax == 1 ? a + 4 * x + 16 * b : a + 4 * b + 16 * x);
}
}  // namespace

template <typename T>
__device__ __forceinline__ Cell<T> sg_load_remote(const SgFusedArgs<T>& A, int rk, int64_t g) {
  if (A.multi)
    return to_cell(A.in_all[0][rk][g], A.in_all[1][rk][g], A.in_all[2][rk][g], A.in_all[3][rk][g], A.in_all[4][rk][g]);
  return to_cell(A.in[0][g], A.in[1][g], A.in[2][g], A.in[3][g], A.in[4][g]);
}

template <typename T, int MINB>
__global__ void __launch_bounds__(256, MINB) sg_fused_kernel(const __grid_constant__ SgFusedArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* cq = reinterpret_cast<T*>(smem_raw);   // [7][G*PS]
  T* fl = cq + NCELLQ * G * PS;             // [3][5][G*NFA]
  __shared__ T ain[G];
  constexpr int CS = G * PS, FS = G * NFA;
  const int     tid = threadIdx.x;
  const int64_t e0  = (int64_t)blockIdx.x * G;
  const int     el  = tid >> 6, c = tid & 63;
  const int     i = c & 3, j = (c >> 2) & 3, k = c >> 4;
  const int64_t e   = e0 + el;
  const bool    on  = e < A.ne;

  // ---- phase 0: own cells (coalesced) ...
  T u[5] = {T(1), T(0), T(0), T(0), T(1)};
  if (on) {
    const int64_t g = e * 64 + c;
#pragma unroll
    for (int v = 0; v < 5; v++) u[v] = A.in[v][g];
    Cell<T> q = to_cell(u[0], u[1], u[2], u[3], u[4]);
    int     s = el * PS + pslot(i + 1, j + 1, k + 1);
    cq[0 * CS + s] = q.rho; cq[1 * CS + s] = q.hx; cq[2 * CS + s] = q.hy; cq[3 * CS + s] = q.hz;
    cq[4 * CS + s] = q.kp;   cq[5 * CS + s] = q.b;  cq[6 * CS + s] = q.q;
    if (c == 0) {
      T edge  = cbrt(A.vol[e]) / T(4);   // kernels.inl:352-354
      ain[el] = edge * edge;
    }
  }
  // ... and the halo layers: 6 sides x 16 cells per element
  for (int it = tid; it < G * 96; it += 256) {
    const int     hel = it / 96, r = it % 96, d = r >> 4, a = r & 3, b = (r >> 2) & 3;
    const int64_t he  = e0 + hel;
    if (he >= A.ne) continue;
    const int ax = d >> 1, up = d & 1;
    const int kd = A.kind[he * 6 + d];
    if (kd == 3) continue;
    Cell<T> q;
    if (kd == 0) {   // wall: mirrored own boundary cell (kernels.inl:169-176)
      int64_t g = he * 64 + cell_ax(ax, up ? 3 : 0, a, b);
      q = to_cell(A.in[0][g], A.in[1][g], A.in[2][g], A.in[3][g], A.in[4][g]);
      q = mirror(q, ax == 0 ? T(1) : T(0), ax == 1 ? T(1) : T(0), ax == 2 ? T(1) : T(0));
    } else {
      int na = a, nb = b;
      if (kd == 2) {
        int qd = A.quad[he * 6 + d];
        na = 2 * (qd & 1) + (a >> 1);
        nb = 2 * (qd >> 1) + (b >> 1);
      }
      int64_t g = (int64_t)A.nid[he * 6 + d] * 64 + cell_ax(ax, up ? 0 : 3, na, nb);
      q = sg_load_remote(A, A.multi ? A.nrk[he * 6 + d] : 0, g);
    }
    int s = hel * PS + pslot_ax(ax, up ? 5 : 0, a, b);
    cq[0 * CS + s] = q.rho; cq[1 * CS + s] = q.hx; cq[2 * CS + s] = q.hy; cq[3 * CS + s] = q.hz;
    cq[4 * CS + s] = q.kp;   cq[5 * CS + s] = q.b;  cq[6 * CS + s] = q.q;
  }
  __syncthreads();

  // ---- phase 1: 3 axes x 5 planes x 16 faces per element, canonical orientation
  for (int it = tid; it < G * 240; it += 256) {
    const int     fel = it / 240, r = it % 240, ax = r / 80, p = (r % 80) >> 4, a = r & 3, b = (r >> 2) & 3;
    const int64_t fe  = e0 + fel;
    if (fe >= A.ne) continue;
    const T nx = ax == 0 ? T(1) : T(0), ny = ax == 1 ? T(1) : T(0), nz = ax == 2 ? T(1) : T(0);
    T F[5];
    T area;
    const int d  = p == 0 ? 2 * ax : 2 * ax + 1;
    const int kd = (p == 0 || p == 4) ? A.kind[fe * 6 + d] : 1;
    if (kd != 3) {
      int sl = fel * PS + pslot_ax(ax, p, a, b), sr = fel * PS + pslot_ax(ax, p + 1, a, b);
      Cell<T> L, R;
      L.rho = cq[0 * CS + sl]; L.hx = cq[1 * CS + sl]; L.hy = cq[2 * CS + sl]; L.hz = cq[3 * CS + sl];
      L.kp   = cq[4 * CS + sl]; L.b  = cq[5 * CS + sl]; L.q  = cq[6 * CS + sl];
      R.rho = cq[0 * CS + sr]; R.hx = cq[1 * CS + sr]; R.hy = cq[2 * CS + sr]; R.hz = cq[3 * CS + sr];
      R.kp   = cq[4 * CS + sr]; R.b  = cq[5 * CS + sr]; R.q  = cq[6 * CS + sr];
      kepes_flux(L, R, nx, ny, nz, F);
      area = (p == 0 || p == 4) ? A.aout[fe * 6 + d] : ain[fel];
    } else {
      // finer neighbours: this coarse boundary cell faces 2x2 fine cells of neighbour quarter (a/2, b/2)
      const int up = p == 4;
      int       so = fel * PS + pslot_ax(ax, up ? 4 : 1, a, b);
      Cell<T> O;
      O.rho = cq[0 * CS + so]; O.hx = cq[1 * CS + so]; O.hy = cq[2 * CS + so]; O.hz = cq[3 * CS + so];
      O.kp   = cq[4 * CS + so]; O.b  = cq[5 * CS + so]; O.q  = cq[6 * CS + so];
      const int row = A.nid[fe * 6 + d], qq = (a >> 1) + 2 * (b >> 1);
      const int64_t nb0 = (int64_t)A.fine_id[row * 4 + qq] * 64;
      const int     rk  = A.multi ? A.fine_rk[row * 4 + qq] : 0;
#pragma unroll
      for (int v = 0; v < 5; v++) F[v] = T(0);
#pragma unroll
      for (int sb = 0; sb < 2; sb++)
#pragma unroll
        for (int sa = 0; sa < 2; sa++) {
          Cell<T> N = sg_load_remote(A, rk, nb0 + cell_ax(ax, up ? 0 : 3, 2 * (a & 1) + sa, 2 * (b & 1) + sb));
          T       Fs[5];
          if (up) kepes_flux(O, N, nx, ny, nz, Fs); else kepes_flux(N, O, nx, ny, nz, Fs);
#pragma unroll
          for (int v = 0; v < 5; v++) F[v] += Fs[v];
        }
      area = A.aout[fe * 6 + d];
    }
    const int fs = fel * NFA + p * 16 + a + 4 * b;
#pragma unroll
    for (int v = 0; v < 5; v++) fl[(ax * 5 + v) * FS + fs] = F[v] * area;
  }

  // phase-2 operands requested before the barrier
  const int stage = A.stage;
  T pv[5] = {u[0], u[1], u[2], u[3], u[4]};
  T vol   = T(1);
  if (on) {
    vol = A.vol[e] / T(64);   // ssp_runge_kutta.inl:116
    if (stage != 1) {
      const int64_t g = e * 64 + c;
#pragma unroll
      for (int v = 0; v < 5; v++) pv[v] = A.prev[v][g];
    }
  }
  __syncthreads();

  // ---- phase 2: each cell gathers its 6 faces (+lower, -upper), RK combination, store
  if (on) {
    const int64_t g  = e * 64 + c;
    const int     bx = el * NFA + j + 4 * k, by = el * NFA + i + 4 * k, bz = el * NFA + i + 4 * j;
    T             sc = rk_scale<T>(stage, A.dt, vol);
#pragma unroll
    for (int v = 0; v < 5; v++) {
      T acc = fl[(0 * 5 + v) * FS + bx + 16 * i] - fl[(0 * 5 + v) * FS + bx + 16 * (i + 1)];
      acc += fl[(1 * 5 + v) * FS + by + 16 * j] - fl[(1 * 5 + v) * FS + by + 16 * (j + 1)];
      acc += fl[(2 * 5 + v) * FS + bz + 16 * k] - fl[(2 * 5 + v) * FS + bz + 16 * (k + 1)];
      A.out[v][g] = rk_apply<T>(stage, pv[v], u[v], acc, sc);
    }
  }
}

template <typename T>
static T* sg_upload(const std::vector<T>& v, int64_t& bytes, cudaError_t& err) {
  T*     d = nullptr;
  size_t n = std::max<size_t>(v.size(), 1) * sizeof(T);
  if (err != cudaSuccess) return nullptr;
  err = cudaMalloc(&d, n);
  if (err != cudaSuccess) return nullptr;
  if (!v.empty()) err = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  bytes += (int64_t)n;
  return d;
}

template <typename T>
static int sg_plan_build(t8b200_subgrid_plan* P, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb,
                         const int32_t* nbr, const T* normals, const T* areas, const int32_t* level_diff,
                         const int32_t* offsets, const int32_t* ranks, const int32_t* indices, int32_t nx,
                         const int32_t* xnbr, const T* xnormals, const T* xareas, const int32_t* xld,
                         const int32_t* xoff) {
  P->ne    = n_local;
  P->multi = n_ghost > 0;
  if (n_ghost > 0 && (!ranks || !indices)) return cudaErrorInvalidValue;
  std::vector<uint8_t> kind(n_local * 6, 255), quad(n_local * 6, 0);
  std::vector<int32_t> nid(n_local * 6, 0), nrk(n_local * 6, 0), fine_id, fine_rk;
  std::vector<T>       aout(n_local * 6, T(0));
  auto owner = [&](int32_t id, int32_t& rk, int32_t& ix) {
    if (id < n_local || !ranks) { rk = ranks ? ranks[id] : 0; ix = ranks ? indices[id] : id; }
    else { rk = ranks[id]; ix = indices[id]; }
  };
  auto face = [&](int32_t l, int32_t r, const T* n, T area, int ld, const int32_t* off) -> int {
    int ax = -1, sg = 0;
    for (int d = 0; d < 3; d++)
      if (n[d] == T(1) || n[d] == T(-1)) { ax = d; sg = n[d] > T(0) ? 1 : 0; }
    if (ax < 0) return cudaErrorInvalidValue;   // the subgrid path is Cartesian-only
    const int dl = 2 * ax + sg, dr = dl ^ 1;
    const int t0 = ax == 0 ? 1 : 0, t1 = ax == 2 ? 1 : 2;   // tangential axes in increasing order
    const T   ca = area / T(16);
    int32_t   rk, ix;
    if (ld == 0) {
      if (l < n_local) { owner(r, rk, ix); kind[l * 6 + dl] = 1; nid[l * 6 + dl] = ix; nrk[l * 6 + dl] = rk; aout[l * 6 + dl] = ca; }
      if (r < n_local) { owner(l, rk, ix); kind[r * 6 + dr] = 1; nid[r * 6 + dr] = ix; nrk[r * 6 + dr] = rk; aout[r * 6 + dr] = ca; }
    } else {   // r is one level coarser than l
      const int qa = off[t0] / 2, qb = off[t1] / 2;
      if (l < n_local) {
        owner(r, rk, ix);
        kind[l * 6 + dl] = 2; nid[l * 6 + dl] = ix; nrk[l * 6 + dl] = rk; aout[l * 6 + dl] = ca;
        quad[l * 6 + dl] = (uint8_t)(qa | (qb << 1));
      }
      if (r < n_local) {
        if (kind[r * 6 + dr] != 3) {
          kind[r * 6 + dr] = 3;
          nid[r * 6 + dr]  = (int32_t)(fine_id.size() / 4);
          fine_id.resize(fine_id.size() + 4, 0);
          fine_rk.resize(fine_rk.size() + 4, 0);
          aout[r * 6 + dr] = ca;
        }
        owner(l, rk, ix);
        fine_id[(size_t)nid[r * 6 + dr] * 4 + qa + 2 * qb] = ix;
        fine_rk[(size_t)nid[r * 6 + dr] * 4 + qa + 2 * qb] = rk;
      }
    }
    return 0;
  };
  for (int32_t f = 0; f < nf; f++) {
    int rc = face(nbr[2 * f], nbr[2 * f + 1], normals + 3 * f, areas[f], level_diff[f], offsets + 3 * f);
    if (rc) return rc;
  }
  for (int32_t f = 0; f < nx; f++) {
    int rc = face(xnbr[2 * f], xnbr[2 * f + 1], xnormals + 3 * f, xareas[f], xld[f], xoff + 3 * f);
    if (rc) return rc;
  }
  for (int32_t b = 0; b < nb; b++) {
    int32_t  e = nbr[2 * (int64_t)nf + b];
    const T* n = normals + 3 * ((int64_t)nf + b);
    int      ax = -1, sg = 0;
    for (int d = 0; d < 3; d++)
      if (n[d] == T(1) || n[d] == T(-1)) { ax = d; sg = n[d] > T(0) ? 1 : 0; }
    if (ax < 0 || e >= n_local) return cudaErrorInvalidValue;
    kind[e * 6 + 2 * ax + sg] = 0;
    aout[e * 6 + 2 * ax + sg] = areas[nf + b] / T(16);
  }
  for (auto kd : kind)
    if (kd == 255) return cudaErrorInvalidValue;   // an element side without a face record: inconsistent connectivity
  cudaError_t err = cudaSuccess;
  P->kind = sg_upload(kind, P->dev_bytes, err);
  P->quad = sg_upload(quad, P->dev_bytes, err);
  P->nid  = sg_upload(nid, P->dev_bytes, err);
  if (P->multi) P->nrk = sg_upload(nrk, P->dev_bytes, err);
  P->aout    = sg_upload(aout, P->dev_bytes, err);
  P->fine_id = sg_upload(fine_id, P->dev_bytes, err);
  if (P->multi) P->fine_rk = sg_upload(fine_rk, P->dev_bytes, err);
  return err;
}

template <typename T>
static int sg_fused_impl(const t8b200_subgrid_plan* P, int stage, const T* const* in, const T* const* const* in_all,
                         const T* const* prev, T* const* out, const T* vol, T dt, void* stream) {
  if (!P || stage < 1 || stage > 3 || !in || !out || !vol || (stage > 1 && !prev)) return cudaErrorInvalidValue;
  if (P->is_f64 != (sizeof(T) == 8) || (P->multi && !in_all)) return cudaErrorInvalidValue;
  if (P->ne == 0) return 0;
  SgFusedArgs<T> A{};
  A.kind = P->kind; A.quad = P->quad; A.nid = P->nid; A.nrk = P->nrk; A.aout = (const T*)P->aout;
  A.fine_id = P->fine_id; A.fine_rk = P->fine_rk;
  for (int k = 0; k < 5; k++) {
    A.in[k] = in[k]; A.in_all[k] = in_all ? in_all[k] : nullptr; A.prev[k] = stage > 1 ? prev[k] : in[k];
    A.out[k] = out[k];
  }
  A.vol = vol; A.dt = dt; A.ne = P->ne; A.stage = stage; A.multi = P->multi;
  constexpr int MINB = sizeof(T) == 8 ? 2 : 4;
  size_t   smem   = sizeof(T) * ((size_t)NCELLQ * G * PS + 15 * (size_t)G * NFA);
  auto     kfn    = sg_fused_kernel<T, MINB>;
  T8B_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned blocks = (unsigned)((P->ne + G - 1) / G);
  kfn<<<blocks, 256, smem, (cudaStream_t)stream>>>(A);
  return cudaGetLastError();
}

extern "C" {

int t8b200_subgrid_plan_create(t8b200_subgrid_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf,
                               int32_t nb, const int32_t* nbr, const void* normals, const void* areas,
                               const int32_t* level_diff, const int32_t* offsets, const int32_t* ranks,
                               const int32_t* indices, int32_t nx, const int32_t* xnbr, const void* xnormals,
                               const void* xareas, const int32_t* xld, const int32_t* xoff) {
  if (!out || n_local < 0 || n_ghost < 0 || nf < 0 || nb < 0 || nx < 0) return cudaErrorInvalidValue;
  if (nf + nb > 0 && (!nbr || !normals || !areas)) return cudaErrorInvalidValue;
  if (nf > 0 && (!level_diff || !offsets)) return cudaErrorInvalidValue;
  if (nx > 0 && (!xnbr || !xnormals || !xareas || !xld || !xoff)) return cudaErrorInvalidValue;
  auto* P   = new t8b200_subgrid_plan();
  P->is_f64 = is_f64 ? 1 : 0;
  int rc = is_f64 ? sg_plan_build<double>(P, n_local, n_ghost, nf, nb, nbr, (const double*)normals,
                                          (const double*)areas, level_diff, offsets, ranks, indices, nx, xnbr,
                                          (const double*)xnormals, (const double*)xareas, xld, xoff)
                  : sg_plan_build<float>(P, n_local, n_ghost, nf, nb, nbr, (const float*)normals, (const float*)areas,
                                         level_diff, offsets, ranks, indices, nx, xnbr, (const float*)xnormals,
                                         (const float*)xareas, xld, xoff);
  if (rc) {
    t8b200_subgrid_plan_destroy(P);
    return rc;
  }
  *out = P;
  return 0;
}
void t8b200_subgrid_plan_destroy(t8b200_subgrid_plan* P) {
  if (!P) return;
  cudaFree(P->kind); cudaFree(P->quad); cudaFree(P->nid); cudaFree(P->nrk); cudaFree(P->aout);
  cudaFree(P->fine_id); cudaFree(P->fine_rk);
  delete P;
}
int t8b200_subgrid_fused_stage_f32(const t8b200_subgrid_plan* plan, int stage, const float* const* in,
                                   const float* const* const* in_all, const float* const* prev, float* const* out,
                                   const float* vol, float dt, void* stream) {
  return sg_fused_impl<float>(plan, stage, in, in_all, prev, out, vol, dt, stream);
}
int t8b200_subgrid_fused_stage_f64(const t8b200_subgrid_plan* plan, int stage, const double* const* in,
                                   const double* const* const* in_all, const double* const* prev, double* const* out,
                                   const double* vol, double dt, void* stream) {
  return sg_fused_impl<double>(plan, stage, in, in_all, prev, out, vol, dt, stream);
}
}
