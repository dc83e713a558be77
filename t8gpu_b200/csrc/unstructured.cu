// Unstructured (MeshManager) hot path for sm_100a: face flux, SSP-RK3 stage, wave-speed reduction, and the fused
// tile-plan stage kernel.  C ABI in include/t8gpu_b200.h.
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
// 2. tile plan + fused stage
// ============================================================================================================

static constexpr int EC = 256;  // elements per chunk == threads per CTA

struct t8b200_plan {
  int     is_f64     = 0;
  int64_t n_local    = 0;
  int     n_chunks   = 0;
  int     max_halo   = 0;
  int     max_faces  = 0;
  int     multi      = 0;  // has ghosts -> needs rank tables
  size_t  smem_bytes = 0;
  int     ms = 0, mf = 0;  // compile-time stride variant selected for the kernel
  int64_t dev_bytes = 0, n_records = 0, n_halo = 0;
  // device arrays
  int32_t*  halo_off  = nullptr;  // n_chunks + 1
  int32_t*  halo_elem = nullptr;  // index into the owner's arrays
  int32_t*  halo_rank = nullptr;  // owner rank (multi only)
  int32_t*  face_off  = nullptr;  // n_chunks + 1
  uint32_t* face_lr   = nullptr;  // slotL | slotR << 16 ; slotR == 0xFFFF -> wall
  void *    fnx = nullptr, *fny = nullptr, *fnz = nullptr, *farea = nullptr;  // general geometry (cmp == 0)
  // compressed geometry (cmp == 1): every normal is +-e_axis and there are <= 256 distinct areas
  int       cmp      = 0;
  int       n_areas  = 0;
  uint2*    face_lg  = nullptr;  // .x = slotL | slotR << 16, .y = axis*2+sign | area index << 3
  void*     area_tab = nullptr;
  int32_t*  csr_base = nullptr;  // n_chunks + 1
  uint16_t* csr_off  = nullptr;  // n_chunks * (EC + 1), relative to csr_base[c]
  uint16_t* csr_ent  = nullptr;  // face_local << 1 | (1 if this element is the right side)
};

template <typename T>
struct FusedArgs {
  const int32_t*  halo_off;
  const int32_t*  halo_elem;
  const int32_t*  halo_rank;
  const int32_t*  face_off;
  const uint32_t* face_lr;
  const T *       fnx, *fny, *fnz, *farea;
  const uint2*    face_lg;
  const T*        area_tab;
  int             n_areas;
  const int32_t*  csr_base;
  const uint16_t* csr_off;
  const uint16_t* csr_ent;
  const T*        in[5];
  const T* const* in_all[5];
  const T*        prev[5];
  T*              out[5];
  const T*        vol;
  T               dt;
  T*              speed_max;
  int64_t         n_local;
  int             stage;
  int             multi;
};

template <typename T, int MS>
__device__ __forceinline__ void store_cell(T* cq, int s, const Cell<T>& q) {
  cq[0 * MS + s] = q.rho; cq[1 * MS + s] = q.vx; cq[2 * MS + s] = q.vy; cq[3 * MS + s] = q.vz;
  cq[4 * MS + s] = q.p;   cq[5 * MS + s] = q.B;  cq[6 * MS + s] = q.w;
}
template <typename T, int MS>
__device__ __forceinline__ Cell<T> load_cell(const T* cq, int s) {
  Cell<T> q;
  q.rho = cq[0 * MS + s]; q.vx = cq[1 * MS + s]; q.vy = cq[2 * MS + s]; q.vz = cq[3 * MS + s];
  q.p   = cq[4 * MS + s]; q.B  = cq[5 * MS + s]; q.w  = cq[6 * MS + s];
  return q;
}

// CTA = one chunk of EC consecutive elements.  MS / MF: compile-time strides of the shared-memory SoA arrays
// (slots = EC own + halo; faces), so every shared access is base + index*sizeof(T) + immediate.
//   phase 0: conserved -> per-cell quantities for the chunk's own elements (coalesced) and its halo (gather)
//   phase 1: every face touching the chunk: flux from the staged cells -> smem (area-scaled)
//   phase 2: per element: signed gather of its faces' fluxes (fixed order: deterministic), RK combination, store
// per-face plan record as the kernel consumes it
template <typename T>
struct FaceRec {
  uint32_t lr;
  T        nx, ny, nz, ar;
};
template <typename T, bool CMP>
__device__ __forceinline__ void load_face_raw(const FusedArgs<T>& A, int g, uint32_t& lr, uint32_t& geo, T& nx, T& ny,
                                              T& nz, T& ar) {
  if (CMP) {
    uint2 v = A.face_lg[g];
    lr  = v.x;
    geo = v.y;
  } else {
    lr = A.face_lr[g];
    nx = A.fnx[g]; ny = A.fny[g]; nz = A.fnz[g]; ar = A.farea[g];
  }
}

template <typename T, int MS, int MF, int MINB, bool CMP>
__global__ void __launch_bounds__(EC, MINB) fused_stage_kernel(const __grid_constant__ FusedArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* cq = reinterpret_cast<T*>(smem_raw);  // [7][MS]
  T* fl = cq + NCELLQ * MS;                // [5][MF]
  __shared__ T atab[CMP ? 256 : 1];
  const int     c   = blockIdx.x;
  const int     tid = threadIdx.x;
  const int64_t e   = (int64_t)c * EC + tid;
  const bool    own = e < A.n_local;

  // ---- phase 0: issue every independent global load of the prologue first
  T u0 = T(1), u1 = T(0), u2 = T(0), u3 = T(0), u4 = T(1);
  if (own) {
    u0 = A.in[0][e]; u1 = A.in[1][e]; u2 = A.in[2][e]; u3 = A.in[3][e]; u4 = A.in[4][e];
  }
  const int h0 = A.halo_off[c], nh = A.halo_off[c + 1] - h0;
  const int f0 = A.face_off[c], nfc = A.face_off[c + 1] - f0;
  // first face record of this thread (consumed in phase 1; its latency hides behind phase 0)
  uint32_t lr_n = 0, geo_n = 0;
  T        nx_n = T(0), ny_n = T(0), nz_n = T(0), ar_n = T(0);
  if (tid < nfc) load_face_raw<T, CMP>(A, f0 + tid, lr_n, geo_n, nx_n, ny_n, nz_n, ar_n);
  if (CMP && tid < A.n_areas) atab[tid] = A.area_tab[tid];

  if (own) store_cell<T, MS>(cq, tid, to_cell(u0, u1, u2, u3, u4));
  for (int h = tid; h < nh; h += EC) {
    int idx = A.halo_elem[h0 + h];
    T   a0, a1, a2, a3, a4;
    if (A.multi) {
      int rk = A.halo_rank[h0 + h];
      a0 = A.in_all[0][rk][idx]; a1 = A.in_all[1][rk][idx]; a2 = A.in_all[2][rk][idx];
      a3 = A.in_all[3][rk][idx]; a4 = A.in_all[4][rk][idx];
    } else {
      a0 = A.in[0][idx]; a1 = A.in[1][idx]; a2 = A.in[2][idx]; a3 = A.in[3][idx]; a4 = A.in[4][idx];
    }
    store_cell<T, MS>(cq, EC + h, to_cell(a0, a1, a2, a3, a4));
  }
  __syncthreads();

  // ---- phase 1: software-pipelined over this thread's faces (record j+EC is in flight while j is evaluated)
  T smax = T(0);
  for (int j = tid; j < nfc; j += EC) {
    uint32_t lr = lr_n, geo = geo_n;
    T        nx = nx_n, ny = ny_n, nz = nz_n, ar = ar_n;
    if (j + EC < nfc) load_face_raw<T, CMP>(A, f0 + j + EC, lr_n, geo_n, nx_n, ny_n, nz_n, ar_n);
    if (CMP) {
      T sg = (geo & 1u) ? T(1) : T(-1);
      int ax = (geo >> 1) & 3;
      nx = ax == 0 ? sg : T(0); ny = ax == 1 ? sg : T(0); nz = ax == 2 ? sg : T(0);
      ar = atab[geo >> 3];
    }
    int     sl = lr & 0xFFFFu, sr = lr >> 16;
    Cell<T> L = load_cell<T, MS>(cq, sl);
    Cell<T> R = sr == 0xFFFF ? mirror(L, nx, ny, nz) : load_cell<T, MS>(cq, sr);
    T F[5];
    T s  = kepes_flux(L, R, nx, ny, nz, F);
    smax = fmax_(smax, s);
#pragma unroll
    for (int k = 0; k < 5; k++) fl[k * MF + j] = ar * F[k];
  }

  // ---- phase 2 operands are requested BEFORE the barrier so their latency overlaps the wait
  const int stage = A.stage;
  int q0 = 0, q1 = 0;
  T   vol = T(1), pv[5] = {u0, u1, u2, u3, u4};
  const uint16_t* ent = A.csr_ent + A.csr_base[c];
  if (own) {
    const uint16_t* off = A.csr_off + (size_t)c * (EC + 1);
    q0  = off[tid];
    q1  = off[tid + 1];
    vol = A.vol[e];
    if (stage != 1) {
#pragma unroll
      for (int k = 0; k < 5; k++) pv[k] = A.prev[k][e];
    }
  }
  __syncthreads();

  if (own) {
    T acc[5] = {T(0), T(0), T(0), T(0), T(0)};
    for (int q = q0; q < q1; q++) {
      int en = ent[q];
      int j  = en >> 1;
      T   sg = (en & 1) ? T(1) : T(-1);
#pragma unroll
      for (int k = 0; k < 5; k++) acc[k] = fma(sg, fl[k * MF + j], acc[k]);
    }
    T sc = rk_scale<T>(stage, A.dt, vol);
    T uin[5] = {u0, u1, u2, u3, u4};
#pragma unroll
    for (int k = 0; k < 5; k++) A.out[k][e] = rk_apply<T>(stage, pv[k], uin[k], acc[k], sc);
  }

  if (A.speed_max) {
    smax = warp_max(smax);
    __shared__ T red[EC / 32];
    if ((tid & 31) == 0) red[tid >> 5] = smax;
    __syncthreads();
    if (tid == 0) {
      T m = red[0];
      for (int w = 1; w < EC / 32; w++) m = fmax_(m, red[w]);
      atomic_max_nonneg(A.speed_max, m);
    }
  }
}

template <typename T>
static T* upload(const std::vector<T>& v, int64_t& bytes, cudaError_t& err) {
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
static int plan_build(t8b200_plan* P, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb, const int32_t* nbr,
                      const T* normals, const T* areas, const int32_t* ranks, const int32_t* indices, int32_t nx,
                      const int32_t* xnbr, const T* xnormals, const T* xareas) {
  const int nchunks = (int)((n_local + EC - 1) / EC);
  P->n_local  = n_local;
  P->n_chunks = nchunks;
  P->multi    = n_ghost > 0;
  if (n_ghost > 0 && (!ranks || !indices)) return cudaErrorInvalidValue;

  // face records: (chunk, global face id) ; global face ids: [0,nf) interior, [nf,nf+nb) boundary, then x-faces
  const int64_t ntot = (int64_t)nf + nb + nx;
  auto endpoints = [&](int64_t f, int32_t& l, int32_t& r) {
    if (f < nf) { l = nbr[2 * f]; r = nbr[2 * f + 1]; }
    else if (f < (int64_t)nf + nb) { l = nbr[2 * (int64_t)nf + (f - nf)]; r = -1; }
    else { int64_t g = f - nf - nb; l = xnbr[2 * g]; r = xnbr[2 * g + 1]; }
  };
  std::vector<int32_t> cnt(nchunks + 1, 0);
  for (int64_t f = 0; f < ntot; f++) {
    int32_t l, r;
    endpoints(f, l, r);
    int cl = l < n_local ? l / EC : -1, cr = (r >= 0 && r < n_local) ? r / EC : -1;
    if (cl < 0 && cr < 0) return cudaErrorInvalidValue;
    if (cl >= 0) cnt[cl + 1]++;
    if (cr >= 0 && cr != cl) cnt[cr + 1]++;
  }
  std::vector<int32_t> face_off(nchunks + 1, 0);
  for (int c = 0; c < nchunks; c++) face_off[c + 1] = face_off[c] + cnt[c + 1];
  const int64_t nrec = face_off[nchunks];
  std::vector<int64_t> rec(nrec);
  {
    std::vector<int32_t> fill(face_off.begin(), face_off.end() - 1);
    for (int64_t f = 0; f < ntot; f++) {
      int32_t l, r;
      endpoints(f, l, r);
      int cl = l < n_local ? l / EC : -1, cr = (r >= 0 && r < n_local) ? r / EC : -1;
      if (cl >= 0) rec[fill[cl]++] = f;
      if (cr >= 0 && cr != cl) rec[fill[cr]++] = f;
    }
  }

  std::vector<int32_t>  halo_off(nchunks + 1, 0), halo_elem, halo_rank, csr_base(nchunks + 1, 0);
  std::vector<uint32_t> face_lr(nrec);
  std::vector<T>        fnx(nrec), fny(nrec), fnz(nrec), far(nrec);
  std::vector<uint16_t> csr_off((size_t)nchunks * (EC + 1)), csr_ent;
  // compressed geometry: all normals +-e_axis, few distinct areas
  std::vector<uint32_t> face_geo(nrec);
  std::vector<T>        area_tab;
  bool                  cmp = true;
  csr_ent.reserve(2 * nrec);
  std::vector<int32_t> halo_tmp, deg(EC + 1);
  int max_halo = 0, max_faces = 0;
  for (int c = 0; c < nchunks; c++) {
    const int64_t e0 = (int64_t)c * EC, e1 = std::min<int64_t>(e0 + EC, n_local);
    const int     r0 = face_off[c], r1 = face_off[c + 1];
    max_faces = std::max(max_faces, r1 - r0);
    if (r1 - r0 > 32767) return cudaErrorInvalidValue;
    // halo = endpoints outside the chunk, sorted + unique
    halo_tmp.clear();
    for (int q = r0; q < r1; q++) {
      int32_t l, r;
      endpoints(rec[q], l, r);
      if (l < e0 || l >= e1) halo_tmp.push_back(l);
      if (r >= 0 && (r < e0 || r >= e1)) halo_tmp.push_back(r);
    }
    std::sort(halo_tmp.begin(), halo_tmp.end());
    halo_tmp.erase(std::unique(halo_tmp.begin(), halo_tmp.end()), halo_tmp.end());
    const int nh = (int)halo_tmp.size();
    if (EC + nh >= 0xFFFF) return cudaErrorInvalidValue;
    max_halo = std::max(max_halo, nh);
    for (int h = 0; h < nh; h++) {
      int32_t id = halo_tmp[h];
      if (id < n_local) {
        halo_elem.push_back(id);
        halo_rank.push_back(ranks ? ranks[id] : 0);
      } else {
        halo_elem.push_back(indices[id]);
        halo_rank.push_back(ranks[id]);
      }
    }
    halo_off[c + 1] = (int32_t)halo_elem.size();
    auto slot_of = [&](int32_t id) -> int {
      if (id >= e0 && id < e1) return (int)(id - e0);
      return EC + (int)(std::lower_bound(halo_tmp.begin(), halo_tmp.end(), id) - halo_tmp.begin());
    };
    // faces + per-element degree
    std::fill(deg.begin(), deg.end(), 0);
    for (int q = r0; q < r1; q++) {
      int64_t f = rec[q];
      int32_t l, r;
      endpoints(f, l, r);
      int sl = slot_of(l), sr = r < 0 ? 0xFFFF : slot_of(r);
      face_lr[q] = (uint32_t)sl | ((uint32_t)sr << 16);
      const T* nrm;
      T        a;
      if (f < (int64_t)nf + nb) { nrm = normals + 3 * f; a = areas[f]; }
      else { int64_t g = f - nf - nb; nrm = xnormals + 3 * g; a = xareas[g]; }
      fnx[q] = nrm[0]; fny[q] = nrm[1]; fnz[q] = nrm[2]; far[q] = a;
      if (cmp) {
        int code = -1;
        for (int d = 0; d < 3; d++) {
          T o1 = nrm[(d + 1) % 3], o2 = nrm[(d + 2) % 3];
          if (o1 == T(0) && o2 == T(0) && (nrm[d] == T(1) || nrm[d] == T(-1))) code = 2 * d + (nrm[d] > T(0) ? 1 : 0);
        }
        int ai = -1;
        for (size_t t = 0; t < area_tab.size(); t++)
          if (area_tab[t] == a) ai = (int)t;
        if (ai < 0 && area_tab.size() < 256) { ai = (int)area_tab.size(); area_tab.push_back(a); }
        if (code < 0 || ai < 0) cmp = false; else face_geo[q] = (uint32_t)code | ((uint32_t)ai << 3);
      }
      if (sl < EC) deg[sl + 1]++;
      if (sr < EC) deg[sr + 1]++;
    }
    for (int i = 0; i < EC; i++) deg[i + 1] += deg[i];
    if (deg[EC] > 65535) return cudaErrorInvalidValue;
    for (int i = 0; i <= EC; i++) csr_off[(size_t)c * (EC + 1) + i] = (uint16_t)deg[i];
    const size_t base = csr_ent.size();
    csr_ent.resize(base + deg[EC]);
    std::vector<int32_t> pos(deg.begin(), deg.end() - 1);
    for (int q = r0; q < r1; q++) {
      uint32_t lr = face_lr[q];
      int      sl = lr & 0xFFFF, sr = lr >> 16, j = q - r0;
      if (sl < EC) csr_ent[base + pos[sl]++] = (uint16_t)(j << 1);
      if (sr < EC) csr_ent[base + pos[sr]++] = (uint16_t)((j << 1) | 1);
    }
    csr_base[c + 1] = (int32_t)csr_ent.size();
  }
  P->max_halo   = max_halo;
  P->max_faces  = max_faces;
  P->n_records  = nrec;
  P->n_halo     = (int64_t)halo_elem.size();
  // stride variants compiled into the library (slots, faces)
  P->ms = EC + max_halo <= 512 ? 512 : 1280;
  P->mf = max_faces <= 1024 ? 1024 : 2560;
  if (EC + max_halo > P->ms || max_faces > P->mf) return cudaErrorInvalidValue;  // chunk too irregular for one CTA
  P->smem_bytes = sizeof(T) * ((size_t)NCELLQ * P->ms + 5 * (size_t)P->mf);
  if (P->smem_bytes > 227 * 1024) return cudaErrorInvalidValue;

  cudaError_t err = cudaSuccess;
  P->halo_off  = upload(halo_off, P->dev_bytes, err);
  P->halo_elem = upload(halo_elem, P->dev_bytes, err);
  if (P->multi) P->halo_rank = upload(halo_rank, P->dev_bytes, err);
  P->face_off = upload(face_off, P->dev_bytes, err);
  P->cmp = cmp ? 1 : 0;
  if (cmp) {
    std::vector<uint2> lg(nrec);
    for (int64_t q = 0; q < nrec; q++) lg[q] = make_uint2(face_lr[q], face_geo[q]);
    P->face_lg  = upload(lg, P->dev_bytes, err);
    P->area_tab = upload(area_tab, P->dev_bytes, err);
    P->n_areas  = (int)area_tab.size();
  } else {
    P->face_lr = upload(face_lr, P->dev_bytes, err);
    P->fnx     = upload(fnx, P->dev_bytes, err);
    P->fny     = upload(fny, P->dev_bytes, err);
    P->fnz     = upload(fnz, P->dev_bytes, err);
    P->farea   = upload(far, P->dev_bytes, err);
  }
  P->csr_base = upload(csr_base, P->dev_bytes, err);
  P->csr_off  = upload(csr_off, P->dev_bytes, err);
  P->csr_ent  = upload(csr_ent, P->dev_bytes, err);
  return err;
}

template <typename T, int MS, int MF, int MINB, bool CMP>
static int launch_variant(const t8b200_plan* P, const FusedArgs<T>& A, cudaStream_t st) {
  auto k = fused_stage_kernel<T, MS, MF, MINB, CMP>;
  T8B_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P->smem_bytes));
  k<<<P->n_chunks, EC, P->smem_bytes, st>>>(A);
  return cudaGetLastError();
}

template <typename T, bool CMP>
static int launch_fused(const t8b200_plan* P, const FusedArgs<T>& A, cudaStream_t st) {
  // resident CTAs per SM are bounded by shared memory; tell ptxas so it can size the register budget
  constexpr int B0 = sizeof(T) == 8 ? 3 : 6;
  if (P->ms == 512 && P->mf == 1024) return launch_variant<T, 512, 1024, B0, CMP>(P, A, st);
  if (P->ms == 512 && P->mf == 2560) return launch_variant<T, 512, 2560, 1, CMP>(P, A, st);
  if (P->ms == 1280 && P->mf == 1024) return launch_variant<T, 1280, 1024, 1, CMP>(P, A, st);
  return launch_variant<T, 1280, 2560, 1, CMP>(P, A, st);
}

template <typename T>
static int fused_stage_impl(const t8b200_plan* P, int stage, const T* const* in, const T* const* const* in_all,
                            const T* const* prev, T* const* out, const T* vol, T dt, T* speed_max, void* stream) {
  if (!P || stage < 1 || stage > 3 || !in || !out || !vol || (stage > 1 && !prev)) return cudaErrorInvalidValue;
  if (P->is_f64 != (sizeof(T) == 8)) return cudaErrorInvalidValue;
  if (P->multi && !in_all) return cudaErrorInvalidValue;
  if (P->n_chunks == 0) return cudaSuccess;
  FusedArgs<T> A{};
  A.halo_off = P->halo_off; A.halo_elem = P->halo_elem; A.halo_rank = P->halo_rank;
  A.face_off = P->face_off; A.face_lr = P->face_lr;
  A.fnx = (const T*)P->fnx; A.fny = (const T*)P->fny; A.fnz = (const T*)P->fnz; A.farea = (const T*)P->farea;
  A.face_lg = P->face_lg; A.area_tab = (const T*)P->area_tab; A.n_areas = P->n_areas;
  A.csr_base = P->csr_base; A.csr_off = P->csr_off; A.csr_ent = P->csr_ent;
  for (int k = 0; k < 5; k++) {
    A.in[k]     = in[k];
    A.in_all[k] = in_all ? in_all[k] : nullptr;
    A.prev[k]   = stage > 1 ? prev[k] : in[k];
    A.out[k]    = out[k];
  }
  A.vol = vol; A.dt = dt; A.speed_max = speed_max; A.n_local = P->n_local;
  A.stage = stage; A.multi = P->multi;
  cudaStream_t st = (cudaStream_t)stream;
  if (speed_max) T8B_TRY(cudaMemsetAsync(speed_max, 0, sizeof(T), st));
  return P->cmp ? launch_fused<T, true>(P, A, st) : launch_fused<T, false>(P, A, st);
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

int t8b200_plan_create(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb,
                       const int32_t* nbr, const void* normals, const void* areas, const int32_t* ranks,
                       const int32_t* indices, int32_t nx, const int32_t* xnbr, const void* xnormals,
                       const void* xareas) {
  if (!out || n_local < 0 || n_ghost < 0 || nf < 0 || nb < 0 || nx < 0) return cudaErrorInvalidValue;
  if ((nf + nb > 0) && (!nbr || !normals || !areas)) return cudaErrorInvalidValue;
  if (nx > 0 && (!xnbr || !xnormals || !xareas)) return cudaErrorInvalidValue;
  t8b200_plan* P = new t8b200_plan();
  P->is_f64      = is_f64 ? 1 : 0;
  int rc = is_f64 ? plan_build<double>(P, n_local, n_ghost, nf, nb, nbr, (const double*)normals, (const double*)areas,
                                       ranks, indices, nx, xnbr, (const double*)xnormals, (const double*)xareas)
                  : plan_build<float>(P, n_local, n_ghost, nf, nb, nbr, (const float*)normals, (const float*)areas,
                                      ranks, indices, nx, xnbr, (const float*)xnormals, (const float*)xareas);
  if (rc != 0) {
    t8b200_plan_destroy(P);
    return rc;
  }
  *out = P;
  return 0;
}

void t8b200_plan_destroy(t8b200_plan* P) {
  if (!P) return;
  cudaFree(P->halo_off); cudaFree(P->halo_elem); cudaFree(P->halo_rank); cudaFree(P->face_off);
  cudaFree(P->face_lr); cudaFree(P->fnx); cudaFree(P->fny); cudaFree(P->fnz); cudaFree(P->farea);
  cudaFree(P->face_lg); cudaFree(P->area_tab);
  cudaFree(P->csr_base); cudaFree(P->csr_off); cudaFree(P->csr_ent);
  delete P;
}

int t8b200_plan_info(const t8b200_plan* P, int64_t info[8]) {
  if (!P || !info) return cudaErrorInvalidValue;
  info[0] = P->n_chunks; info[1] = P->max_halo; info[2] = P->max_faces; info[3] = (int64_t)P->smem_bytes;
  info[4] = P->dev_bytes; info[5] = P->n_records; info[6] = P->n_halo; info[7] = EC;
  return 0;
}

int t8b200_fused_stage_f32(const t8b200_plan* plan, int stage, const float* const* in,
                           const float* const* const* in_all, const float* const* prev, float* const* out,
                           const float* vol, float dt, float* speed_max_dev, void* stream) {
  return fused_stage_impl<float>(plan, stage, in, in_all, prev, out, vol, dt, speed_max_dev, stream);
}
int t8b200_fused_stage_f64(const t8b200_plan* plan, int stage, const double* const* in,
                           const double* const* const* in_all, const double* const* prev, double* const* out,
                           const double* vol, double dt, double* speed_max_dev, void* stream) {
  return fused_stage_impl<double>(plan, stage, in, in_all, prev, out, vol, dt, speed_max_dev, stream);
}

}  // extern "C"
