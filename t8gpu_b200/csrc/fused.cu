// Fused tile-plan RK stage for unstructured (MeshManager) meshes on sm_100a: face flux + per-element accumulation +
// SSP-RK3 combination + wave-speed reduction in ONE kernel per stage; the flux accumulators never reach HBM.
// C ABI in include/t8gpu_b200.h (section 2).
//
// Reference behaviour replaced (not translated): one stage of CompressibleEulerSolver::iterate,
//   examples/compressible_euler/solver.cu:78-112 = kepes_compute_fluxes (kernels.cu:135-309) +
//   reflective_boundary_condition (kernels.cu:311-469) + SSP_3RK_step{1,2,3} (ssp_runge_kutta.inl:30-99),
//   and the thrust::reduce of solver.cu:213-217.
//
// Plan ("tile plan", built once per connectivity by t8b200_plan_create): per chunk of EC = 256 consecutive elements
//   header   8 x int32: -, -, nh | nfc << 16, e0 | e1 << 16, e2, ovf_off_base, ovf_ent_base, area
//   halo     sorted unique elements outside the chunk that share a face with it (slot EC + h); fixed stride HS per
//            chunk (padding = -1), so the indices can be requested without waiting for the header
//   faces    one 32-bit record slotL | slotR << 16 per face touching the chunk, fixed stride FS per chunk.
//            Cartesian forests ("cmp": every normal +-e_axis, <= 256 distinct areas): records are put in canonical
//            orientation (normal = +e_axis, sides swapped where the stored normal was -e_axis) and grouped by axis,
//            [0,e0) x, [e0,e1) y, [e1,e2) z, then wall faces [e2,nfc) with the outward normal coded in the slotR field;
//            a chunk whose faces all have the same area carries it in the header (applied once per element).
//            General meshes: normals and areas as four T arrays, wall = slotR 0xFFFF.
//   ell      per element 8 x uint16 entries (face_local << 1 | side), 0xFFFF = none; one 128-bit load per thread.
//            Elements with more than 8 faces (hanging faces on several sides) continue in a per-chunk overflow CSR.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../include/t8gpu_b200.h"
#include "common.cuh"
#include "euler_flux.cuh"

using namespace t8b200;

static constexpr int EC  = 256;  // elements per chunk == threads per CTA
static constexpr int ELL = 8;    // face entries per element held in the fixed-width table

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
  int     cmp = 0, n_areas = 0;
  int     hs = 0, fs = 0;  // per-chunk strides of the halo and face arrays
  // device arrays
  int32_t*  hdr       = nullptr;  // 8 per chunk
  int32_t*  halo_elem = nullptr;  // index into the owner's arrays
  int32_t*  halo_rank = nullptr;  // owner rank (multi only)
  uint32_t* face_lr   = nullptr;
  uint8_t*  face_ai   = nullptr;  // cmp: area index per record (read only by chunks with mixed areas)
  void *    fnx = nullptr, *fny = nullptr, *fnz = nullptr, *farea = nullptr;  // general geometry (cmp == 0)
  void*     area_tab = nullptr;
  uint4*    ell      = nullptr;  // n_chunks * EC
  uint16_t* ovf_off  = nullptr;  // (EC + 1) per chunk that has overflow entries
  uint16_t* ovf_ent  = nullptr;
};

template <typename T>
struct FusedArgs {
  const int4*     hdr;
  const int32_t*  halo_elem;
  const int32_t*  halo_rank;
  const uint32_t* face_lr;
  const uint8_t*  face_ai;
  const T *       fnx, *fny, *fnz, *farea;
  const T*        area_tab;
  int             n_areas;
  int             hs, fs;
  const uint4*    ell;
  const uint16_t* ovf_off;
  const uint16_t* ovf_ent;
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
  cq[0 * MS + s] = q.rho; cq[1 * MS + s] = q.hx; cq[2 * MS + s] = q.hy; cq[3 * MS + s] = q.hz;
  cq[4 * MS + s] = q.kp;  cq[5 * MS + s] = q.b;  cq[6 * MS + s] = q.q;
}
template <typename T, int MS>
__device__ __forceinline__ Cell<T> load_cell(const T* cq, int s) {
  Cell<T> q;
  q.rho = cq[0 * MS + s]; q.hx = cq[1 * MS + s]; q.hy = cq[2 * MS + s]; q.hz = cq[3 * MS + s];
  q.kp  = cq[4 * MS + s]; q.b  = cq[5 * MS + s]; q.q  = cq[6 * MS + s];
  return q;
}

// one entry of the element -> face table: acc += (side ? +1 : -1) * flux[face]
template <typename T, int MF>
__device__ __forceinline__ void gather_entry(const T* fl, unsigned en, T acc[5]) {
  const int j  = en >> 1;
  const T   sg = (en & 1u) ? T(1) : T(-1);
#pragma unroll
  for (int k = 0; k < 5; k++) acc[k] = fma(sg, fl[k * MF + j], acc[k]);
}

// bulk L2 prefetch of [p, p + bytes): 16-byte granules, address aligned down (a hint; used on the state rows inside
// [0, n_local) and on plan arrays, which upload() pads)
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
  const unsigned long long a = (unsigned long long)p & ~15ull;
  bytes = (bytes + (unsigned)((unsigned long long)p & 15ull)) & ~15u;
  if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes) : "memory");
}

// CTA = one chunk of EC consecutive elements.  MS / MF: compile-time strides of the shared-memory SoA arrays
// (slots = EC own + halo; faces), so every shared access is base + index*sizeof(T) + immediate.
//   phase 0: conserved -> per-cell quantities for the chunk's own elements (coalesced) and its halo (gather)
//   phase 1: every face touching the chunk: flux from the staged cells -> smem
//   phase 2: per element: signed gather of its faces' fluxes (fixed order: deterministic), RK combination, store
template <typename T, int MS, int MF, int MINB, bool CMP>
__global__ void __launch_bounds__(EC, MINB) fused_stage_kernel(const __grid_constant__ FusedArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* cq = reinterpret_cast<T*>(smem_raw);  // [7][MS]
  T* fl = cq + NCELLQ * MS;                // [5][MF]
  __shared__ T atab[CMP ? 256 : 1];
  __shared__ T red[EC / 32];
  const int     c   = blockIdx.x;
  const int     tid = threadIdx.x;
  const int64_t e   = (int64_t)c * EC + tid;
  const bool    own = e < A.n_local;

  // ---- phase 0: issue every independent global load of the prologue first.  The halo indices and the first face
  // record sit at fixed strides, so nothing here waits for the chunk header.
  const int4 h0v = __ldg(A.hdr + 2 * c), h1v = __ldg(A.hdr + 2 * c + 1);
  const int64_t hb = (int64_t)c * A.hs, fb = (int64_t)c * A.fs;
  T   u0 = T(1), u1 = T(0), u2 = T(0), u3 = T(0), u4 = T(1);
  int hidx = -1, hrk = 0;
  if (tid < A.hs) {
    hidx = A.halo_elem[hb + tid];
    if (A.multi) hrk = A.halo_rank[hb + tid];
  }
  if (own) {
    u0 = A.in[0][e]; u1 = A.in[1][e]; u2 = A.in[2][e]; u3 = A.in[3][e]; u4 = A.in[4][e];
  }
  // first face record of this thread (consumed in phase 1; its latency hides behind phase 0)
  uint32_t lr_n = 0;
  if (tid < A.fs) lr_n = A.face_lr[fb + tid];
  if (tid < 8) {  // phase-2 operands of this chunk -> L2 now, so that their loads before the barrier are L2 hits
    const int64_t  b0 = (int64_t)c * EC;
    const unsigned n0 = (unsigned)min((int64_t)EC, A.n_local - b0);
    if (tid < 5) { if (A.stage != 1) prefetch_l2(A.prev[tid] + b0, n0 * sizeof(T)); }
    else if (tid == 5) prefetch_l2(A.vol + b0, n0 * sizeof(T));
    else if (tid == 6) prefetch_l2(A.ell + b0, n0 * sizeof(uint4));
  }
  if (CMP && tid < A.n_areas) atab[tid] = A.area_tab[tid];

  if (own) store_cell<T, MS>(cq, tid, to_cell(u0, u1, u2, u3, u4));
  for (int h = tid; h < A.hs; h += EC) {
    if (h >= EC) {  // only plans with more than EC halo entries in some chunk (adaptive meshes)
      hidx = A.halo_elem[hb + h];
      if (A.multi) hrk = A.halo_rank[hb + h];
    }
    if (hidx >= 0) {
      T a0, a1, a2, a3, a4;
      if (A.multi) {
        a0 = A.in_all[0][hrk][hidx]; a1 = A.in_all[1][hrk][hidx]; a2 = A.in_all[2][hrk][hidx];
        a3 = A.in_all[3][hrk][hidx]; a4 = A.in_all[4][hrk][hidx];
      } else {
        a0 = A.in[0][hidx]; a1 = A.in[1][hidx]; a2 = A.in[2][hidx]; a3 = A.in[3][hidx]; a4 = A.in[4][hidx];
      }
      store_cell<T, MS>(cq, EC + h, to_cell(a0, a1, a2, a3, a4));
    }
  }
  const int  nfc = (unsigned)h0v.z >> 16;
  const int  e0 = h0v.w & 0xFFFF, e1 = (unsigned)h0v.w >> 16, e2 = h1v.x;
  const int  area_idx  = h1v.w;
  const bool want_smax = A.speed_max != nullptr;
  __syncthreads();

  // ---- phase 1: software-pipelined over this thread's faces (record j+EC is in flight while j is evaluated)
  T smax = T(0);
  if (CMP) {
    // SCALE = false: every face of the chunk has the same area, applied once per element in phase 2
    auto interior = [&](auto scale_tag) {
      constexpr bool SCALE = decltype(scale_tag)::value;
      for (int j = tid; j < e2; j += EC) {
        const uint32_t lr = lr_n;
        if (j + EC < nfc) lr_n = A.face_lr[fb + j + EC];
        const Cell<T> L = load_cell<T, MS>(cq, lr & 0xFFFFu);
        const Cell<T> R = load_cell<T, MS>(cq, lr >> 16);
        T F[5], s;
        if (j < e0) s = kepes_flux_n<T, 0>(L, R, T(0), T(0), T(0), F);
        else if (j < e1) s = kepes_flux_n<T, 1>(L, R, T(0), T(0), T(0), F);
        else s = kepes_flux_n<T, 2>(L, R, T(0), T(0), T(0), F);
        if (want_smax) smax = fmax_(smax, s);
        if (SCALE) {
          const T ar = atab[A.face_ai[fb + j]];
#pragma unroll
          for (int k = 0; k < 5; k++) F[k] *= ar;
        }
#pragma unroll
        for (int k = 0; k < 5; k++) fl[k * MF + j] = F[k];
      }
    };
    if (area_idx >= 0) interior(std::false_type{}); else interior(std::true_type{});
    // wall faces: slotR field = 0xFFF8 | (axis << 1 | sign of the outward normal)
    for (int j = e2 + tid; j < nfc; j += EC) {
      const uint32_t lr   = A.face_lr[fb + j];
      const int      code = (lr >> 16) & 7;
      const T        sg   = (code & 1) ? T(1) : T(-1);
      const int      ax   = code >> 1;
      const T nx = ax == 0 ? sg : T(0), ny = ax == 1 ? sg : T(0), nz = ax == 2 ? sg : T(0);
      const Cell<T> L = load_cell<T, MS>(cq, lr & 0xFFFFu);
      const Cell<T> R = mirror(L, nx, ny, nz);
      T F[5];
      smax = fmax_(smax, kepes_flux_n<T, -1>(L, R, nx, ny, nz, F));
      if (area_idx < 0) {
        const T ar = atab[A.face_ai[fb + j]];
#pragma unroll
        for (int k = 0; k < 5; k++) F[k] *= ar;
      }
#pragma unroll
      for (int k = 0; k < 5; k++) fl[k * MF + j] = F[k];
    }
  } else {
    T nx_n = T(0), ny_n = T(0), nz_n = T(0), ar_n = T(0);
    if (tid < nfc) { nx_n = A.fnx[fb + tid]; ny_n = A.fny[fb + tid]; nz_n = A.fnz[fb + tid]; ar_n = A.farea[fb + tid]; }
    for (int j = tid; j < nfc; j += EC) {
      const uint32_t lr = lr_n;
      const T        nx = nx_n, ny = ny_n, nz = nz_n, ar = ar_n;
      if (j + EC < nfc) {
        const int64_t g = fb + j + EC;
        lr_n = A.face_lr[g]; nx_n = A.fnx[g]; ny_n = A.fny[g]; nz_n = A.fnz[g]; ar_n = A.farea[g];
      }
      const int     sr = lr >> 16;
      const Cell<T> L  = load_cell<T, MS>(cq, lr & 0xFFFFu);
      const Cell<T> R  = sr == 0xFFFF ? mirror(L, nx, ny, nz) : load_cell<T, MS>(cq, sr);
      T F[5];
      smax = fmax_(smax, kepes_flux_n<T, -1>(L, R, nx, ny, nz, F));
#pragma unroll
      for (int k = 0; k < 5; k++) fl[k * MF + j] = ar * F[k];
    }
  }

  // ---- phase 2: the element->face table, the volume and the old states are requested BEFORE the barrier so their
  // latency overlaps the wait.  The element's own conserved values are read again (L1/L2 hit) instead of being
  // carried in 10 registers through the face loop, and combined with U^n right away (5 live values, not 10).
  const int stage = A.stage;
  uint4     el = make_uint4(~0u, ~0u, ~0u, ~0u);
  T         vol = T(1), base[5] = {T(0), T(0), T(0), T(0), T(0)};
  if (own) {
    el  = A.ell[e];
    vol = A.vol[e];
#pragma unroll
    for (int k = 0; k < 5; k++) base[k] = A.in[k][e];
    if (stage != 1) {
      const T cp = stage == 2 ? T(0.75) : T(0.33333333333333), ci = stage == 2 ? T(0.25) : T(0.66666666666666);
#pragma unroll
      for (int k = 0; k < 5; k++) base[k] = cp * A.prev[k][e] + ci * base[k];
    }
  }
  __syncthreads();

  if (own) {
    T acc[5] = {T(0), T(0), T(0), T(0), T(0)};
    const unsigned w[4] = {el.x, el.y, el.z, el.w};
#pragma unroll
    for (int s = 0; s < ELL; s++) {
      const unsigned en = (s & 1) ? w[s >> 1] >> 16 : w[s >> 1] & 0xFFFFu;
      if (en != 0xFFFFu) gather_entry<T, MF>(fl, en, acc);
    }
    if (h1v.y >= 0) {  // rare: elements of this chunk with more than ELL faces
      const uint16_t* off = A.ovf_off + h1v.y;
      const uint16_t* ent = A.ovf_ent + h1v.z;
#pragma unroll 1
      for (int q = off[tid], q1 = off[tid + 1]; q < q1; q++) gather_entry<T, MF>(fl, ent[q], acc);
    }
    T sc = fast_rcp(vol) * A.dt;
    if (stage == 2) sc *= T(0.25);
    if (stage == 3) sc *= T(0.66666666666666);
    if (CMP && area_idx >= 0) sc *= atab[area_idx];
#pragma unroll
    for (int k = 0; k < 5; k++) A.out[k][e] = base[k] + sc * acc[k];
  }

  if (A.speed_max) {
    smax = warp_max(smax);
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
  size_t n = std::max<size_t>(v.size(), 1) * sizeof(T) + 32;  // slack for 16-byte granular prefetch hints
  if (err != cudaSuccess) return nullptr;
  err = cudaMalloc(&d, n);
  if (err != cudaSuccess) return nullptr;
  if (!v.empty()) err = cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  bytes += (int64_t)n;
  return d;
}

// axis-aligned unit normal -> axis * 2 + (1 if positive), else -1
template <typename T>
static int axis_code(const T* n) {
  for (int d = 0; d < 3; d++) {
    T o1 = n[(d + 1) % 3], o2 = n[(d + 2) % 3];
    if (o1 == T(0) && o2 == T(0) && (n[d] == T(1) || n[d] == T(-1))) return 2 * d + (n[d] > T(0) ? 1 : 0);
  }
  return -1;
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

  // global face ids: [0,nf) interior, [nf,nf+nb) boundary, then the extra partition-boundary faces
  const int64_t ntot = (int64_t)nf + nb + nx;
  auto endpoints = [&](int64_t f, int32_t& l, int32_t& r) {
    if (f < nf) { l = nbr[2 * f]; r = nbr[2 * f + 1]; }
    else if (f < (int64_t)nf + nb) { l = nbr[2 * (int64_t)nf + (f - nf)]; r = -1; }
    else { int64_t g = f - nf - nb; l = xnbr[2 * g]; r = xnbr[2 * g + 1]; }
  };
  auto geometry = [&](int64_t f, const T*& nrm, T& a) {
    if (f < (int64_t)nf + nb) { nrm = normals + 3 * f; a = areas[f]; }
    else { int64_t g = f - nf - nb; nrm = xnormals + 3 * g; a = xareas[g]; }
  };

  // compressed geometry possible?  (every normal +-e_axis, <= 256 distinct areas)
  bool                 cmp = true;
  std::vector<T>       area_tab;
  std::vector<uint8_t> area_of(ntot);
  for (int64_t f = 0; f < ntot && cmp; f++) {
    const T* nrm;
    T        a;
    geometry(f, nrm, a);
    if (axis_code(nrm) < 0) { cmp = false; break; }
    int ai = -1;
    for (size_t t = area_tab.size(); t-- > 0;)
      if (area_tab[t] == a) { ai = (int)t; break; }
    if (ai < 0) {
      if (area_tab.size() >= 256) { cmp = false; break; }
      ai = (int)area_tab.size();
      area_tab.push_back(a);
    }
    area_of[f] = (uint8_t)ai;
  }

  // bucket faces by chunk (a face between two chunks appears in both)
  std::vector<int32_t> cnt(nchunks + 1, 0);
  for (int64_t f = 0; f < ntot; f++) {
    int32_t l, r;
    endpoints(f, l, r);
    int cl = l < n_local ? l / EC : -1, cr = (r >= 0 && r < n_local) ? r / EC : -1;
    if (cl < 0 && cr < 0) return cudaErrorInvalidValue;
    if (cl >= 0) cnt[cl + 1]++;
    if (cr >= 0 && cr != cl) cnt[cr + 1]++;
  }
  std::vector<int64_t> face_off(nchunks + 1, 0);
  for (int c = 0; c < nchunks; c++) face_off[c + 1] = face_off[c] + cnt[c + 1];
  const int64_t nrec = face_off[nchunks];
  if (nrec > 0x7FFFFFFF) return cudaErrorInvalidValue;
  std::vector<int64_t> rec(nrec);
  {
    std::vector<int64_t> fill(face_off.begin(), face_off.end() - 1);
    for (int64_t f = 0; f < ntot; f++) {
      int32_t l, r;
      endpoints(f, l, r);
      int cl = l < n_local ? l / EC : -1, cr = (r >= 0 && r < n_local) ? r / EC : -1;
      if (cl >= 0) rec[fill[cl]++] = f;
      if (cr >= 0 && cr != cl) rec[fill[cr]++] = f;
    }
  }

  int max_faces = 0;
  for (int c = 0; c < nchunks; c++) max_faces = std::max(max_faces, cnt[c + 1]);
  if (max_faces > 32767) return cudaErrorInvalidValue;
  const int     FS   = (max_faces + 31) / 32 * 32;  // per-chunk stride of the face arrays
  const int64_t nfix = (int64_t)nchunks * FS;
  std::vector<int32_t>  hdr((size_t)nchunks * 8, 0), halo_elem, halo_rank, halo_cnt(nchunks, 0);
  std::vector<uint32_t> face_lr(nfix, 0);
  std::vector<uint8_t>  face_ai(cmp ? nfix : 0, 0);
  std::vector<T>        fnx(cmp ? 0 : nfix), fny(cmp ? 0 : nfix), fnz(cmp ? 0 : nfix), far(cmp ? 0 : nfix);
  std::vector<uint16_t> ell((size_t)nchunks * EC * ELL, 0xFFFF), ovf_off, ovf_ent;
  std::vector<int32_t>  halo_tmp;
  std::vector<int64_t>  order;   // records of the chunk in kernel order
  std::vector<std::vector<uint16_t>> per_el(EC);
  int max_halo = 0;
  for (int c = 0; c < nchunks; c++) {
    const int64_t e0 = (int64_t)c * EC, e1 = std::min<int64_t>(e0 + EC, n_local);
    const int64_t r0 = face_off[c], r1 = face_off[c + 1];
    const int     nfc = (int)(r1 - r0);
    // halo = endpoints outside the chunk, sorted + unique
    halo_tmp.clear();
    for (int64_t q = r0; q < r1; q++) {
      int32_t l, r;
      endpoints(rec[q], l, r);
      if (l < e0 || l >= e1) halo_tmp.push_back(l);
      if (r >= 0 && (r < e0 || r >= e1)) halo_tmp.push_back(r);
    }
    std::sort(halo_tmp.begin(), halo_tmp.end());
    halo_tmp.erase(std::unique(halo_tmp.begin(), halo_tmp.end()), halo_tmp.end());
    const int nh = (int)halo_tmp.size();
    if (EC + nh >= 0xFFF0) return cudaErrorInvalidValue;
    max_halo = std::max(max_halo, nh);
    int32_t* H = &hdr[(size_t)c * 8];
    halo_cnt[c] = nh;
    H[2] = nh | (nfc << 16);
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
    auto slot_of = [&](int32_t id) -> int {
      if (id >= e0 && id < e1) return (int)(id - e0);
      return EC + (int)(std::lower_bound(halo_tmp.begin(), halo_tmp.end(), id) - halo_tmp.begin());
    };
    // kernel order of the records: cmp -> x, y, z interior faces, then walls; general -> as enumerated
    order.assign(rec.begin() + r0, rec.begin() + r1);
    int seg[4] = {0, 0, 0, 0};
    if (cmp) {
      auto key = [&](int64_t f) -> int {
        int32_t l, r;
        endpoints(f, l, r);
        if (r < 0) return 3;
        const T* nrm;
        T        a;
        geometry(f, nrm, a);
        return axis_code(nrm) >> 1;
      };
      std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key(a) < key(b); });
      for (int64_t f : order) seg[key(f)]++;
    }
    H[3] = seg[0] | ((seg[0] + seg[1]) << 16);
    H[4] = seg[0] + seg[1] + seg[2];
    for (auto& v : per_el) v.clear();
    int  area0 = -1;
    bool uniform = cmp;
    for (int j = 0; j < nfc; j++) {
      const int64_t f = order[j], q = (int64_t)c * FS + j;
      int32_t       l, r;
      endpoints(f, l, r);
      const T* nrm;
      T        a;
      geometry(f, nrm, a);
      int sl = slot_of(l), sr = r < 0 ? 0xFFFF : slot_of(r);
      if (cmp) {
        const int code = axis_code(nrm);
        if (r < 0) sr = 0xFFF8 | code;                  // wall: outward normal coded in the slotR field
        else if (!(code & 1)) std::swap(sl, sr);        // canonical orientation: normal = +e_axis
        face_ai[q] = area_of[f];
        if (area0 < 0) area0 = area_of[f];
        if (area_of[f] != area0) uniform = false;
      } else {
        fnx[q] = nrm[0]; fny[q] = nrm[1]; fnz[q] = nrm[2]; far[q] = a;
      }
      face_lr[q] = (uint32_t)sl | ((uint32_t)sr << 16);
      if (sl < EC) per_el[sl].push_back((uint16_t)(j << 1));
      if (sr < EC) per_el[sr].push_back((uint16_t)((j << 1) | 1));
    }
    H[7] = (uniform && area0 >= 0) ? area0 : -1;
    // fixed-width table + overflow CSR
    bool overflow = false;
    for (int i = 0; i < EC; i++) {
      const auto& v = per_el[i];
      for (size_t s = 0; s < v.size() && s < (size_t)ELL; s++) ell[((size_t)c * EC + i) * ELL + s] = v[s];
      if (v.size() > (size_t)ELL) overflow = true;
    }
    H[5] = -1;
    H[6] = 0;
    if (overflow) {
      if (ovf_off.size() + EC + 1 > 0x7FFFFFFF || ovf_ent.size() > 0x7FFFFFFF) return cudaErrorInvalidValue;
      H[5] = (int32_t)ovf_off.size();
      H[6] = (int32_t)ovf_ent.size();
      size_t n = 0;
      for (int i = 0; i < EC; i++) {
        ovf_off.push_back((uint16_t)n);
        for (size_t s = ELL; s < per_el[i].size(); s++) { ovf_ent.push_back(per_el[i][s]); n++; }
      }
      if (n > 65535) return cudaErrorInvalidValue;
      ovf_off.push_back((uint16_t)n);
    }
  }
  // halo lists at a fixed stride per chunk, padded with -1
  const int HS = std::max(32, (max_halo + 31) / 32 * 32);
  {
    std::vector<int32_t> he((size_t)nchunks * HS, -1), hr(P->multi ? (size_t)nchunks * HS : 0, 0);
    size_t src = 0;
    for (int c = 0; c < nchunks; c++)
      for (int h = 0; h < halo_cnt[c]; h++, src++) {
        he[(size_t)c * HS + h] = halo_elem[src];
        if (P->multi) hr[(size_t)c * HS + h] = halo_rank[src];
      }
    P->n_halo = (int64_t)halo_elem.size();
    halo_elem.swap(he);
    halo_rank.swap(hr);
  }
  P->hs = HS;
  P->fs = FS;
  P->max_halo   = max_halo;
  P->max_faces  = max_faces;
  P->n_records  = nrec;
  // stride variants compiled into the library (slots, faces)
  P->ms = EC + max_halo <= 512 ? 512 : 1280;
  P->mf = max_faces <= 1024 ? 1024 : 2560;
  if (EC + max_halo > P->ms || max_faces > P->mf) return cudaErrorInvalidValue;  // chunk too irregular for one CTA
  P->smem_bytes = sizeof(T) * ((size_t)NCELLQ * P->ms + 5 * (size_t)P->mf);
  if (P->smem_bytes > 220 * 1024) return cudaErrorInvalidValue;

  cudaError_t err = cudaSuccess;
  P->hdr       = upload(hdr, P->dev_bytes, err);
  P->halo_elem = upload(halo_elem, P->dev_bytes, err);
  if (P->multi) P->halo_rank = upload(halo_rank, P->dev_bytes, err);
  P->face_lr = upload(face_lr, P->dev_bytes, err);
  P->cmp     = cmp ? 1 : 0;
  if (cmp) {
    P->face_ai  = upload(face_ai, P->dev_bytes, err);
    P->area_tab = upload(area_tab, P->dev_bytes, err);
    P->n_areas  = (int)area_tab.size();
  } else {
    P->fnx   = upload(fnx, P->dev_bytes, err);
    P->fny   = upload(fny, P->dev_bytes, err);
    P->fnz   = upload(fnz, P->dev_bytes, err);
    P->farea = upload(far, P->dev_bytes, err);
  }
  P->ell     = reinterpret_cast<uint4*>(upload(ell, P->dev_bytes, err));
  P->ovf_off = upload(ovf_off, P->dev_bytes, err);
  P->ovf_ent = upload(ovf_ent, P->dev_bytes, err);
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
  constexpr int B0 = sizeof(T) == 8 ? 3 : 5;
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
  A.hdr = reinterpret_cast<const int4*>(P->hdr);
  A.halo_elem = P->halo_elem; A.halo_rank = P->halo_rank;
  A.face_lr = P->face_lr; A.face_ai = P->face_ai;
  A.fnx = (const T*)P->fnx; A.fny = (const T*)P->fny; A.fnz = (const T*)P->fnz; A.farea = (const T*)P->farea;
  A.area_tab = (const T*)P->area_tab; A.n_areas = P->n_areas; A.hs = P->hs; A.fs = P->fs;
  A.ell = P->ell; A.ovf_off = P->ovf_off; A.ovf_ent = P->ovf_ent;
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
  cudaFree(P->hdr); cudaFree(P->halo_elem); cudaFree(P->halo_rank);
  cudaFree(P->face_lr); cudaFree(P->face_ai); cudaFree(P->fnx); cudaFree(P->fny); cudaFree(P->fnz); cudaFree(P->farea);
  cudaFree(P->area_tab); cudaFree(P->ell); cudaFree(P->ovf_off); cudaFree(P->ovf_ent);
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
