// Device-side connectivity for ADAPTIVE (2:1 face-balanced) Cartesian forests (SURVEY f-2, second half): from the
// leaves of a one-tree quad / hex forest -- Morton key of the anchor and level of every leaf, in SFC order, as t8code
// enumerates them -- to the arrays MeshManager::compute_connectivity_information uploads
// (t8gpu/mesh/mesh_manager.inl:332-481), bit for bit: hanging faces emitted from the fine side, ghost faces first and
// owned by the lower rank, boundary faces behind the interior ones; plus the x-faces (faces whose ghost belongs to a
// lower rank) the owner-computes scheme needs.  Replaces the reference's serial host loop over
// t8_forest_leaf_face_neighbors (one heap allocation per face) by a binary search per face neighbour on the device:
// the leaf that contains a point is the last leaf whose key is not larger than the point's key.
//
// t8code semantics assumed as in cartesian.cu: Morton order with x = bit 0, face ids -x,+x,-y,+y,-z,+z,
// finer neighbours in child-id order, contiguous SFC partition with first element of rank p = floor(N p / P), ghosts in
// SFC order.  Keys are at resolution 2^-20 per axis (interleaved, dim bits per level).
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include "../../include/t8gpu_b200.h"
#include "common.cuh"

namespace {

constexpr int MAXL = 20;

struct Forest {
  int             dim, periodic, nranks, rank;
  int64_t         N, lo, hi;
  const uint64_t* key;
  const int32_t*  level;
};

__host__ __device__ inline int64_t part_off(int64_t N, int P, int p) {
  return (int64_t)(((unsigned __int128)N * (unsigned)p) / (unsigned)P);
}
__device__ inline int owner_of(const Forest& f, int64_t e) {
  int p = (int)(((unsigned __int128)e * (unsigned)f.nranks) / (unsigned __int128)f.N);
  while (p + 1 < f.nranks && part_off(f.N, f.nranks, p + 1) <= e) p++;
  while (p > 0 && part_off(f.N, f.nranks, p) > e) p--;
  return p;
}
__device__ inline uint32_t compact_bits(uint64_t k, int dim) {
  uint32_t r = 0;
  for (int b = 0; b < MAXL; b++) r |= (uint32_t)((k >> (dim * b)) & 1u) << b;
  return r;
}
__device__ inline uint64_t spread_bits(uint32_t v, int dim) {
  uint64_t r = 0;
  for (int b = 0; b < MAXL; b++) r |= (uint64_t)((v >> b) & 1u) << (dim * b);
  return r;
}
__device__ inline uint64_t morton(const uint32_t c[3], int dim) {
  uint64_t k = 0;
  for (int d = 0; d < dim; d++) k |= spread_bits(c[d], dim) << d;
  return k;
}
// the leaf containing key k: last leaf with key <= k
__device__ inline int64_t find_leaf(const Forest& f, uint64_t k) {
  int64_t lo = 0, hi = f.N - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (f.key[mid] <= k) lo = mid; else hi = mid - 1;
  }
  return lo;
}
// t8_forest_leaf_face_neighbors: 0 (domain boundary), 1 (same size or coarser) or 2^(dim-1) (finer, child-id order)
__device__ inline int face_neighbors(const Forest& f, int64_t e, int face, int64_t out[4]) {
  const int dim = f.dim, l = f.level[e], ax = face >> 1, up = face & 1;
  uint32_t  c[3] = {0, 0, 0};
  for (int d = 0; d < dim; d++) c[d] = compact_bits(f.key[e] >> d, dim);
  const uint32_t h = 1u << (MAXL - l), full = 1u << MAXL;
  int64_t        nc = (int64_t)c[ax] + (up ? (int64_t)h : -(int64_t)h);
  if (nc < 0 || nc >= (int64_t)full) {
    if (!f.periodic) return 0;
    nc = (nc + full) % full;
  }
  uint32_t a[3] = {c[0], c[1], c[2]};
  a[ax]         = (uint32_t)nc;
  const int64_t idx = find_leaf(f, morton(a, dim));
  if (f.level[idx] <= l) { out[0] = idx; return 1; }
  const uint32_t hh = h >> 1;
  int            nn = 0;
  for (int ch = 0; ch < (1 << dim); ch++) {
    if (((ch >> ax) & 1) != (up ? 0 : 1)) continue;   // the children of the same-size neighbour that touch the face
    uint32_t b[3] = {a[0], a[1], a[2]};
    for (int d = 0; d < dim; d++)
      if ((ch >> d) & 1) b[d] += hh;
    out[nn++] = find_leaf(f, morton(b, dim));
  }
  return nn;
}

__global__ void ghost_candidates_kernel(Forest f, unsigned long long* counter, int64_t* cand) {
  const int64_t e = f.lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= f.hi) return;
  for (int face = 0; face < 2 * f.dim; face++) {
    int64_t   nb[4];
    const int nn = face_neighbors(f, e, face, nb);
    for (int i = 0; i < nn; i++)
      if (nb[i] < f.lo || nb[i] >= f.hi) {
        const unsigned long long pos = atomicAdd(counter, 1ull);
        if (cand) cand[pos] = nb[i];
      }
  }
}

__device__ inline int64_t ghost_pos(const int64_t* ghosts, int64_t ng, int64_t key) {
  int64_t lo = 0, hi = ng - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ghosts[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct FaceOut {   // interior / owned faces and x-faces; ld / off only for the subgrid layout
  int32_t *nbr, *xnbr, *ld, *off, *xld, *xoff;
  void *   normals, *areas, *xnormals, *xareas;
};

// MODE 0: faces per element (interior, boundary, x);  MODE 1: fill at the scanned offsets.  SUB: the layout of
// SubgridMeshManager (subgrid_mesh_manager.inl:560-786 add_face): dim normal components, level difference and
// neighbour offset per face, the finer element second (normal flipped when the pair is swapped).
template <typename T, int MODE, bool SUB>
__global__ void faces_kernel(Forest f, const int64_t* __restrict__ ghosts, int64_t ng, int64_t* cnt, int64_t* bcnt,
                             int64_t* xcnt, int64_t nf_total, FaceOut out) {
  const int64_t nl = f.hi - f.lo, ei = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ei >= nl) return;
  const int64_t e = f.lo + ei;
  const int     l = f.level[e], nd = SUB ? f.dim : 3, E = 4;
  const double  h = ldexp(1.0, -l), area = f.dim == 3 ? h * h : h;
  int64_t o = MODE ? cnt[ei] : 0, bo = MODE ? bcnt[ei] : 0, xo = MODE ? xcnt[ei] : 0;
  auto emit = [&](bool x, int64_t at, int64_t n_global, int32_t nid, int face, double a) {
    int32_t* pn = x ? out.xnbr : out.nbr;
    T*       pr = (T*)(x ? out.xnormals : out.normals);
    T*       pa = (T*)(x ? out.xareas : out.areas);
    T        sgn = (face & 1) ? T(1) : T(-1);
    int32_t  left = (int32_t)ei, right = nid;
    if (SUB) {
      int32_t*  pl = x ? out.xld : out.ld;
      int32_t*  po = x ? out.xoff : out.off;
      const int nlev = f.level[n_global], ax = face >> 1;
      int       ofs[3] = {0, 0, 0}, ldiff = 0;
      if (nlev == l) {
        ofs[ax] = (face & 1) ? 0 : E - 1;
      } else if (nlev < l) {   // coarser neighbour: which quadrant of its face this element touches
        const int child = (int)((f.key[e] >> (f.dim * (MAXL - l))) & ((1u << f.dim) - 1));
        for (int d = 0; d < f.dim; d++) ofs[d] = E / 2 * ((child >> d) & 1);
        ofs[ax] = (face & 1) ? 0 : E - 1;
        ldiff   = nlev - l;
      } else {                 // finer neighbour: swap so that the finer element comes second
        const int child = (int)((f.key[n_global] >> (f.dim * (MAXL - nlev))) & ((1u << f.dim) - 1));
        for (int d = 0; d < f.dim; d++) ofs[d] = E / 2 * ((child >> d) & 1);
        ofs[ax] = (face & 1) ? E - 1 : 0;
        ldiff   = l - nlev;
        left = nid; right = (int32_t)ei; sgn = -sgn;
      }
      pl[at] = ldiff;
      for (int d = 0; d < f.dim; d++) po[f.dim * at + d] = ofs[d];
    }
    pn[2 * at] = left; pn[2 * at + 1] = right;
    for (int d = 0; d < nd; d++) pr[nd * at + d] = T(0);
    pr[nd * at + (face >> 1)] = sgn;
    pa[at] = (T)a;
  };
  for (int face = 0; face < 2 * f.dim; face++) {
    int64_t   nb[4];
    const int nn = face_neighbors(f, e, face, nb);
    // ghost neighbours first (mesh_manager.inl:396-409, subgrid_mesh_manager.inl:859-877): the lower rank owns the
    // face, area / num_neighbors
    for (int i = 0; i < nn; i++) {
      if (nb[i] >= f.lo && nb[i] < f.hi) continue;
      const int32_t nid = (int32_t)(nl + ghost_pos(ghosts, ng, nb[i]));
      if (f.rank < owner_of(f, nb[i])) {
        if (MODE) emit(false, o, nb[i], nid, face, area / (double)nn);
        o++;
      } else {
        if (MODE) emit(true, xo, nb[i], nid, face, area / (double)nn);
        xo++;
      }
    }
    // local neighbour (mesh_manager.inl:411-424): once per pair, hanging faces from the fine side
    if (nn == 1 && nb[0] >= f.lo && nb[0] < f.hi) {
      const int64_t nid = nb[0] - f.lo;
      if (nid > ei || (nid < ei && f.level[nb[0]] < l)) {
        if (MODE) emit(false, o, nb[0], (int32_t)nid, face, area);
        o++;
      }
    }
    if (nn == 0) {   // domain boundary (mesh_manager.inl:431-440): behind the interior faces
      if (MODE) {
        const int64_t at = nf_total + bo;
        T*            pr = (T*)out.normals;
        out.nbr[2 * nf_total + bo] = (int32_t)ei;
        for (int d = 0; d < nd; d++) pr[nd * at + d] = T(0);
        pr[nd * at + (face >> 1)] = (face & 1) ? T(1) : T(-1);
        ((T*)out.areas)[at] = (T)area;
      }
      bo++;
    }
  }
  if (!MODE) { cnt[ei] = o; bcnt[ei] = bo; xcnt[ei] = xo; }
}

template <typename T>
__global__ void elements_kernel(Forest f, const int64_t* __restrict__ ghosts, int64_t ng, int32_t* ranks,
                                int32_t* indices, T* volumes, T* centroids) {
  const int64_t nl = f.hi - f.lo, i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nl) {
    const int64_t e = f.lo + i;
    const double  h = ldexp(1.0, -f.level[e]);
    ranks[i]   = f.rank;
    indices[i] = (int32_t)i;
    volumes[i] = (T)(f.dim == 3 ? h * h * h : h * h);
    for (int d = 0; d < 3; d++)
      centroids[3 * i + d] = d < f.dim ? (T)(ldexp((double)compact_bits(f.key[e] >> d, f.dim), -MAXL) + 0.5 * h) : T(0);
  } else if (i < nl + ng) {
    const int64_t n = ghosts[i - nl];
    const int     p = owner_of(f, n);
    ranks[i]   = p;
    indices[i] = (int32_t)(n - part_off(f.N, f.nranks, p));
  }
}

template <typename T, bool SUB>
int build(t8b200_cart_conn* out, t8b200_subgrid_face_info* info, Forest f, cudaStream_t st) {
  const int64_t  nl     = f.hi - f.lo;
  const unsigned blocks = (unsigned)((nl + 255) / 256);
  auto           pol    = thrust::cuda::par.on(st);
  T8bScratch<int64_t> ghosts_mem, cnt_mem;
  int64_t*&           ghosts = ghosts_mem.p;
  int64_t             ng     = 0;
  if (f.nranks > 1 && nl > 0) {
    T8bScratch<unsigned long long> counter_mem;
    unsigned long long*&           counter = counter_mem.p;
    T8B_TRY(cudaMalloc(&counter, 8));
    T8B_TRY(cudaMemsetAsync(counter, 0, 8, st));
    ghost_candidates_kernel<<<blocks, 256, 0, st>>>(f, counter, nullptr);
    unsigned long long nc = 0;
    T8B_TRY(cudaMemcpyAsync(&nc, counter, 8, cudaMemcpyDeviceToHost, st));
    T8B_TRY(cudaStreamSynchronize(st));
    if (nc > 0) {
      T8B_TRY(cudaMalloc(&ghosts, nc * sizeof(int64_t)));
      T8B_TRY(cudaMemsetAsync(counter, 0, 8, st));
      ghost_candidates_kernel<<<blocks, 256, 0, st>>>(f, counter, ghosts);
      thrust::sort(pol, ghosts, ghosts + nc);
      ng = thrust::unique(pol, ghosts, ghosts + nc) - ghosts;
    }
  }
  int64_t*& cnt = cnt_mem.p;   // three count arrays of nl + 1 entries
  T8B_TRY(cudaMalloc(&cnt, 3 * (nl + 1) * sizeof(int64_t)));
  T8B_TRY(cudaMemsetAsync(cnt, 0, 3 * (nl + 1) * sizeof(int64_t), st));
  int64_t *bcnt = cnt + (nl + 1), *xcnt = cnt + 2 * (nl + 1);
  if (nl > 0)
    faces_kernel<T, 0, SUB><<<blocks, 256, 0, st>>>(f, ghosts, ng, cnt, bcnt, xcnt, 0, FaceOut{});
  thrust::exclusive_scan(pol, cnt, cnt + nl + 1, cnt);
  thrust::exclusive_scan(pol, bcnt, bcnt + nl + 1, bcnt);
  thrust::exclusive_scan(pol, xcnt, xcnt + nl + 1, xcnt);
  int64_t nf = 0, nb = 0, nx = 0;
  T8B_TRY(cudaMemcpyAsync(&nf, cnt + nl, 8, cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaMemcpyAsync(&nb, bcnt + nl, 8, cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaMemcpyAsync(&nx, xcnt + nl, 8, cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaStreamSynchronize(st));
  if (nf + nb > 0x7FFFFF00LL || nl + ng > 0x7FFFFF00LL) return cudaErrorInvalidValue;
  out->n_local = nl; out->n_ghost = ng; out->n_faces = nf; out->n_bfaces = nb; out->n_xfaces = nx;
  auto alloc = [](void** p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 16); };
  T8B_TRY(alloc((void**)&out->ranks, (nl + ng) * 4));
  T8B_TRY(alloc((void**)&out->indices, (nl + ng) * 4));
  T8B_TRY(alloc((void**)&out->face_neighbors, (2 * nf + nb) * 4));
  const int nd = SUB ? f.dim : 3;
  T8B_TRY(alloc(&out->face_normals, (nf + nb) * nd * sizeof(T)));
  T8B_TRY(alloc(&out->face_surfaces, (nf + nb) * sizeof(T)));
  T8B_TRY(alloc(&out->volumes, nl * sizeof(T)));
  T8B_TRY(alloc(&out->centroids, nl * 3 * sizeof(T)));
  T8B_TRY(alloc((void**)&out->x_face_neighbors, nx * 8));
  T8B_TRY(alloc(&out->x_face_normals, nx * nd * sizeof(T)));
  T8B_TRY(alloc(&out->x_face_surfaces, nx * sizeof(T)));
  FaceOut fo{out->face_neighbors, out->x_face_neighbors, nullptr, nullptr, nullptr, nullptr,
             out->face_normals,   out->face_surfaces,    out->x_face_normals, out->x_face_surfaces};
  if (SUB) {
    T8B_TRY(alloc((void**)&info->level_diff, nf * 4));
    T8B_TRY(alloc((void**)&info->offsets, nf * f.dim * 4));
    T8B_TRY(alloc((void**)&info->x_level_diff, nx * 4));
    T8B_TRY(alloc((void**)&info->x_offsets, nx * f.dim * 4));
    fo.ld = info->level_diff; fo.off = info->offsets; fo.xld = info->x_level_diff; fo.xoff = info->x_offsets;
  }
  if (nl > 0) faces_kernel<T, 1, SUB><<<blocks, 256, 0, st>>>(f, ghosts, ng, cnt, bcnt, xcnt, nf, fo);
  if (nl + ng > 0)
    elements_kernel<T><<<(unsigned)((nl + ng + 255) / 256), 256, 0, st>>>(f, ghosts, ng, out->ranks, out->indices,
                                                                           (T*)out->volumes, (T*)out->centroids);
  T8B_TRY(cudaGetLastError());
  T8B_TRY(cudaStreamSynchronize(st));
  return 0;
}

// on failure nothing stays allocated behind `out` / `info`
template <typename T, bool SUB>
int build_or_release(t8b200_cart_conn* out, t8b200_subgrid_face_info* info, Forest f, cudaStream_t st) {
  int rc;
  try {
    rc = build<T, SUB>(out, info, f, st);
  } catch (const std::bad_alloc&) {
    rc = cudaErrorMemoryAllocation;
  } catch (...) {
    cudaGetLastError();
    rc = cudaErrorUnknown;
  }
  if (rc) {
    t8b200_cartesian_connectivity_free(out);
    if (info) t8b200_subgrid_face_info_free(info);
  }
  return rc;
}

}  // namespace

extern "C" int t8b200_forest_connectivity(t8b200_cart_conn* out, int is_f64, int dim, int periodic, int64_t n_leaves,
                                          const uint64_t* keys_dev, const int32_t* levels_dev, int nranks, int rank,
                                          void* stream) {
  if (!out || (dim != 2 && dim != 3) || n_leaves < 0 || nranks < 1 || rank < 0 || rank >= nranks) return cudaErrorInvalidValue;
  if (n_leaves > 0 && (!keys_dev || !levels_dev)) return cudaErrorInvalidValue;
  *out = t8b200_cart_conn{};
  Forest f{dim, periodic ? 1 : 0, nranks, rank, n_leaves, part_off(n_leaves, nranks, rank),
           part_off(n_leaves, nranks, rank + 1), keys_dev, levels_dev};
  return is_f64 ? build_or_release<double, false>(out, nullptr, f, (cudaStream_t)stream)
                : build_or_release<float, false>(out, nullptr, f, (cudaStream_t)stream);
}

extern "C" int t8b200_forest_subgrid_connectivity(t8b200_cart_conn* out, t8b200_subgrid_face_info* info, int is_f64,
                                                  int dim, int periodic, int64_t n_leaves, const uint64_t* keys_dev,
                                                  const int32_t* levels_dev, int nranks, int rank, void* stream) {
  if (!out || !info || (dim != 2 && dim != 3) || n_leaves < 0 || nranks < 1 || rank < 0 || rank >= nranks)
    return cudaErrorInvalidValue;
  if (n_leaves > 0 && (!keys_dev || !levels_dev)) return cudaErrorInvalidValue;
  *out  = t8b200_cart_conn{};
  *info = t8b200_subgrid_face_info{};
  Forest f{dim, periodic ? 1 : 0, nranks, rank, n_leaves, part_off(n_leaves, nranks, rank),
           part_off(n_leaves, nranks, rank + 1), keys_dev, levels_dev};
  return is_f64 ? build_or_release<double, true>(out, info, f, (cudaStream_t)stream)
                : build_or_release<float, true>(out, info, f, (cudaStream_t)stream);
}

extern "C" void t8b200_subgrid_face_info_free(t8b200_subgrid_face_info* info) {
  if (!info) return;
  cudaFree(info->level_diff); cudaFree(info->offsets); cudaFree(info->x_level_diff); cudaFree(info->x_offsets);
  *info = t8b200_subgrid_face_info{};
}
