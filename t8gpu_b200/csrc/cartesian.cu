// Device-side connectivity for uniform periodic Cartesian forests and the Kelvin-Helmholtz initial state.
// Produces the arrays of MeshManager::compute_connectivity_information (t8gpu/mesh/mesh_manager.inl:332-481) for
// t8_cmesh_new_periodic(dim) + t8_forest_new_uniform(level) on rank `rank` of `nranks`, without the serial host
// loop over t8_forest_leaf_face_neighbors (one heap allocation per face in the reference).
// t8code semantics assumed: Morton order with x = bit 0, face ids -x,+x,-y,+y,-z,+z, contiguous SFC partition with
// first element of rank p = floor(N p / P), ghosts ordered by SFC index.
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/scan.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include <cmath>

#include "../../include/t8gpu_b200.h"
#include "common.cuh"

namespace {

struct Geo {
  int     dim, level, nranks, rank;
  int     b[3];       // trees per axis ("brick" of unit trees, periodic over the whole brick); tree id = x + bx (y + by z)
  int64_t N, lo, hi;  // N = global element count; elements are ordered by tree, Morton inside a tree
};

__host__ __device__ inline int64_t part_off(int64_t N, int P, int p) {
  return (int64_t)(((unsigned __int128)N * (unsigned)p) / (unsigned)P);
}

__device__ inline int owner_of(const Geo& g, int64_t e) {
  int p = (int)(((unsigned __int128)e * (unsigned)g.nranks) / (unsigned __int128)g.N);
  while (p + 1 < g.nranks && part_off(g.N, g.nranks, p + 1) <= e) p++;
  while (p > 0 && part_off(g.N, g.nranks, p) > e) p--;
  return p;
}

__device__ inline void decode(int64_t m, int dim, int level, int c[3]) {
  c[0] = c[1] = c[2] = 0;
  for (int b = 0; b < level; b++)
    for (int d = 0; d < dim; d++) c[d] |= (int)((m >> (dim * b + d)) & 1) << b;
}
__device__ inline int64_t encode(const int c[3], int dim, int level) {
  int64_t m = 0;
  for (int b = 0; b < level; b++)
    for (int d = 0; d < dim; d++) m |= (int64_t)((c[d] >> b) & 1) << (dim * b + d);
  return m;
}
// global element id -> brick-global cell coordinates
__device__ inline void gdecode(const Geo& g, int64_t m, int c[3]) {
  const int64_t per = (int64_t)1 << (g.dim * g.level);
  int64_t       t   = m / per;
  decode(m - t * per, g.dim, g.level, c);
  int tc[3] = {(int)(t % g.b[0]), (int)((t / g.b[0]) % g.b[1]), (int)(t / ((int64_t)g.b[0] * g.b[1]))};
  for (int d = 0; d < g.dim; d++) c[d] += tc[d] << g.level;
}
__device__ inline int64_t gencode(const Geo& g, const int c[3]) {
  const int ext = 1 << g.level;
  int       tc[3] = {0, 0, 0}, lc[3] = {0, 0, 0};
  for (int d = 0; d < g.dim; d++) { tc[d] = c[d] >> g.level; lc[d] = c[d] & (ext - 1); }
  int64_t t = tc[0] + (int64_t)g.b[0] * (tc[1] + (int64_t)g.b[1] * tc[2]);
  return (t << (g.dim * g.level)) + encode(lc, g.dim, g.level);
}
__device__ inline int64_t neighbor(const Geo& g, const int c[3], int face) {
  int n[3] = {c[0], c[1], c[2]};
  int ax = face >> 1, ext = g.b[ax] << g.level;
  n[ax] += (face & 1) ? 1 : -1;
  if (n[ax] < 0) n[ax] += ext;
  if (n[ax] >= ext) n[ax] -= ext;
  return gencode(g, n);
}

__global__ void count_ghost_candidates(Geo g, unsigned long long* counter, int64_t* cand) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.hi - g.lo) return;
  int c[3];
  gdecode(g, g.lo + e, c);
  for (int f = 0; f < 2 * g.dim; f++) {
    int64_t n = neighbor(g, c, f);
    if (n < g.lo || n >= g.hi) {
      unsigned long long pos = atomicAdd(counter, 1ull);
      if (cand) cand[pos] = n;
    }
  }
}

__device__ inline int32_t ghost_pos(const int64_t* ghosts, int64_t ng, int64_t key) {
  int64_t lo = 0, hi = ng - 1;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (ghosts[mid] < key) lo = mid + 1; else hi = mid;
  }
  return (int32_t)lo;
}

// mode 0: count faces per element (cnt main, xcnt extra); mode 1: fill using the scanned offsets
template <typename T, int MODE>
__global__ void faces_kernel(Geo g, const int64_t* __restrict__ ghosts, int64_t ng, int64_t* cnt, int64_t* xcnt,
                             int32_t* nbr, T* normals, T* areas, int32_t* xnbr, T* xnormals, T* xareas) {
  int64_t nl = g.hi - g.lo;
  int64_t e  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nl) return;
  int c[3];
  gdecode(g, g.lo + e, c);
  int64_t o = MODE ? cnt[e] : 0, xo = MODE ? xcnt[e] : 0;
  double  h    = ldexp(1.0, -g.level);
  T       area = (T)(g.dim == 3 ? h * h : h);
  for (int f = 0; f < 2 * g.dim; f++) {
    int64_t n = neighbor(g, c, f);
    bool    main = false, extra = false;
    int32_t nid = 0;
    if (n >= g.lo && n < g.hi) {
      nid  = (int32_t)(n - g.lo);
      main = nid > e;  // mesh_manager.inl:411-414 (uniform: the coarser-neighbour clause never fires)
    } else {
      nid   = (int32_t)(nl + ghost_pos(ghosts, ng, n));
      main  = g.rank < owner_of(g, n);  // mesh_manager.inl:397
      extra = !main;
    }
    if (main) {
      if (MODE) {
        nbr[2 * o] = (int32_t)e; nbr[2 * o + 1] = nid;
        normals[3 * o] = T(0); normals[3 * o + 1] = T(0); normals[3 * o + 2] = T(0);
        normals[3 * o + (f >> 1)] = (f & 1) ? T(1) : T(-1);
        areas[o] = area;
      }
      o++;
    } else if (extra) {
      if (MODE) {
        xnbr[2 * xo] = (int32_t)e; xnbr[2 * xo + 1] = nid;
        xnormals[3 * xo] = T(0); xnormals[3 * xo + 1] = T(0); xnormals[3 * xo + 2] = T(0);
        xnormals[3 * xo + (f >> 1)] = (f & 1) ? T(1) : T(-1);
        xareas[xo] = area;
      }
      xo++;
    }
  }
  if (!MODE) { cnt[e] = o; xcnt[e] = xo; }
}

template <typename T>
__global__ void elements_kernel(Geo g, const int64_t* __restrict__ ghosts, int64_t ng, int32_t* ranks,
                                int32_t* indices, T* volumes, T* centroids) {
  int64_t nl = g.hi - g.lo;
  int64_t i  = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nl) {
    int c[3];
    gdecode(g, g.lo + i, c);
    double h = ldexp(1.0, -g.level);
    ranks[i]   = g.rank;
    indices[i] = (int32_t)i;
    volumes[i] = (T)(g.dim == 3 ? h * h * h : h * h);
    for (int d = 0; d < 3; d++) centroids[3 * i + d] = d < g.dim ? (T)((c[d] + 0.5) * h) : T(0);
  } else if (i < nl + ng) {
    int64_t n = ghosts[i - nl];
    int     p = owner_of(g, n);
    ranks[i]   = p;
    indices[i] = (int32_t)(n - part_off(g.N, g.nranks, p));
  }
}

// Cartesian Kelvin-Helmholtz field; mixed-precision literals follow examples/subgrid/solver.inl:36-56 / 82-103.
template <typename T>
struct Ptrs5 { T* p[5]; };

template <typename T>
__global__ void kh_kernel(int dim, int64_t n, const T* __restrict__ centers, Ptrs5<T> u) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T x = centers[3 * i], s = dim == 3 ? centers[3 * i + 2] : centers[3 * i + 1];
  T gamma = T(1.4);
  T sigma = (T)(0.05f / sqrtf(2.0f));
  bool inside = fabs((double)s - 0.5) < 0.25;
  T rho = (T)(inside ? 2.0 : 1.0);
  T m1  = (T)(inside ? -0.5 : 0.5);
  T a = (s - 0.75f) / (2 * sigma);
  T b = (s - 0.25f) / (2 * sigma);
  T g = exp(-a * a) + exp(-b * b);
  double pert = 0.1 * sin(4.0f * M_PI * ((double)x - 0.5)) * (double)g;
  T mp = (T)((double)rho * pert);
  T m2 = dim == 3 ? T(0) : mp, m3 = dim == 3 ? mp : T(0);
  u.p[0][i] = rho; u.p[1][i] = m1; u.p[2][i] = m2; u.p[3][i] = m3;
  u.p[4][i] = T(2.5) / (gamma - T(1.0)) + T(0.5) * (m1 * m1 + m2 * m2 + m3 * m3) / rho;
}

// The Kelvin-Helmholtz state on the globe of the unstructured example (examples/compressible_euler/solver.cu:17-72: a
// host lambda per element in the reference) at n element centroids, same expression types: float_type square roots and
// inverse trigonometric functions, double for the terms that involve the double literals (:54-61).
template <typename T>
__global__ void spherical_kh_kernel(int64_t n, const T* __restrict__ centers, Ptrs5<T> u) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T sigma = T(0.2) / sqrt(T(2.0)), gamma = T(1.4);
  const T x = centers[3 * i], y = centers[3 * i + 1], z = centers[3 * i + 2];
  const T r = sqrt(x * x + y * y + z * z);
  const T er0 = x / r, er1 = y / r, er2 = z / r;
  const T ep0 = er1 / sqrt(er1 * er1 + er0 * er0), ep1 = -er0 / sqrt(er1 * er1 + er0 * er0), ep2 = T(0.0);
  const T et0 = er1 * ep2 - er2 * ep1, et1 = er2 * ep0 - er0 * ep2, et2 = er0 * ep1 - er1 * ep0;
  const T phi   = (T)((y >= 0.0) ? (double)acos(x / sqrt(x * x + y * y)) : 2.0 * M_PI - acos(x / sqrt(x * x + y * y)));
  const T theta = asin(z / r);
  const T v_phi   = (T)(r * cos(theta) * (theta < 0 ? -0.5 : 0.5));
  const T v_theta = (T)(0.5 * r * sin(2.0 * phi) * (exp(-(theta / (2 * sigma)) * (theta / (2 * sigma)))));
  const T rho = (T)(theta < 0.0 ? 2.0 : 1.0);
  const T m1 = rho * (v_phi * ep0 + v_theta * et0), m2 = rho * (v_phi * ep1 + v_theta * et1),
          m3 = rho * (v_phi * ep2 + v_theta * et2);
  u.p[0][i] = rho; u.p[1][i] = m1; u.p[2][i] = m2; u.p[3][i] = m3;
  u.p[4][i] = T(2.5) / (gamma - T(1.0)) + T(0.5) * (m1 * m1 + m2 * m2 + m3 * m3) / rho;
}

template <typename T>
int build(t8b200_cart_conn* out, Geo g, cudaStream_t st) {
  const int64_t nl = g.hi - g.lo;
  const unsigned blocks = (unsigned)((nl + 255) / 256);
  auto pol = thrust::cuda::par.on(st);
  int64_t* ghosts = nullptr;
  int64_t  ng     = 0;
  if (g.nranks > 1 && nl > 0) {
    unsigned long long* counter;
    T8B_TRY(cudaMalloc(&counter, 8));
    T8B_TRY(cudaMemsetAsync(counter, 0, 8, st));
    count_ghost_candidates<<<blocks, 256, 0, st>>>(g, counter, nullptr);
    unsigned long long nc = 0;
    T8B_TRY(cudaMemcpyAsync(&nc, counter, 8, cudaMemcpyDeviceToHost, st));
    T8B_TRY(cudaStreamSynchronize(st));
    if (nc > 0) {
      T8B_TRY(cudaMalloc(&ghosts, nc * sizeof(int64_t)));
      T8B_TRY(cudaMemsetAsync(counter, 0, 8, st));
      count_ghost_candidates<<<blocks, 256, 0, st>>>(g, counter, ghosts);
      thrust::sort(pol, ghosts, ghosts + nc);
      ng = thrust::unique(pol, ghosts, ghosts + nc) - ghosts;
    }
    cudaFree(counter);
  }
  int64_t *cnt, *xcnt;
  T8B_TRY(cudaMalloc(&cnt, (nl + 1) * sizeof(int64_t)));
  T8B_TRY(cudaMalloc(&xcnt, (nl + 1) * sizeof(int64_t)));
  T8B_TRY(cudaMemsetAsync(cnt, 0, (nl + 1) * sizeof(int64_t), st));
  T8B_TRY(cudaMemsetAsync(xcnt, 0, (nl + 1) * sizeof(int64_t), st));
  if (nl > 0)
    faces_kernel<T, 0><<<blocks, 256, 0, st>>>(g, ghosts, ng, cnt, xcnt, nullptr, nullptr, nullptr, nullptr, nullptr,
                                               nullptr);
  thrust::exclusive_scan(pol, cnt, cnt + nl + 1, cnt);
  thrust::exclusive_scan(pol, xcnt, xcnt + nl + 1, xcnt);
  int64_t nf = 0, nx = 0;
  T8B_TRY(cudaMemcpyAsync(&nf, cnt + nl, 8, cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaMemcpyAsync(&nx, xcnt + nl, 8, cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaStreamSynchronize(st));
  out->n_local = nl; out->n_ghost = ng; out->n_faces = nf; out->n_bfaces = 0; out->n_xfaces = nx;
  auto alloc = [](void** p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 16); };
  T8B_TRY(alloc((void**)&out->ranks, (nl + ng) * 4));
  T8B_TRY(alloc((void**)&out->indices, (nl + ng) * 4));
  T8B_TRY(alloc((void**)&out->face_neighbors, nf * 8));
  T8B_TRY(alloc(&out->face_normals, nf * 3 * sizeof(T)));
  T8B_TRY(alloc(&out->face_surfaces, nf * sizeof(T)));
  T8B_TRY(alloc(&out->volumes, nl * sizeof(T)));
  T8B_TRY(alloc(&out->centroids, nl * 3 * sizeof(T)));
  T8B_TRY(alloc((void**)&out->x_face_neighbors, nx * 8));
  T8B_TRY(alloc(&out->x_face_normals, nx * 3 * sizeof(T)));
  T8B_TRY(alloc(&out->x_face_surfaces, nx * sizeof(T)));
  if (nl > 0)
    faces_kernel<T, 1><<<blocks, 256, 0, st>>>(g, ghosts, ng, cnt, xcnt, out->face_neighbors, (T*)out->face_normals,
                                               (T*)out->face_surfaces, out->x_face_neighbors, (T*)out->x_face_normals,
                                               (T*)out->x_face_surfaces);
  if (nl + ng > 0)
    elements_kernel<T><<<(unsigned)((nl + ng + 255) / 256), 256, 0, st>>>(g, ghosts, ng, out->ranks, out->indices,
                                                                           (T*)out->volumes, (T*)out->centroids);
  T8B_TRY(cudaGetLastError());
  T8B_TRY(cudaStreamSynchronize(st));
  cudaFree(cnt); cudaFree(xcnt); cudaFree(ghosts);
  return 0;
}

template <typename T>
int kh_impl(int dim, int64_t n, const T* centers, T* const* u, void* stream) {
  if ((dim != 2 && dim != 3) || n < 0 || !centers || !u) return cudaErrorInvalidValue;
  if (n == 0) return 0;
  Ptrs5<T> p;
  for (int k = 0; k < 5; k++) p.p[k] = u[k];
  kh_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dim, n, centers, p);
  return cudaGetLastError();
}

template <typename T>
int spherical_kh_impl(int64_t n, const T* centers, T* const* u, void* stream) {
  if (n < 0 || (n > 0 && (!centers || !u))) return cudaErrorInvalidValue;
  if (n == 0) return 0;
  Ptrs5<T> p;
  for (int k = 0; k < 5; k++) {
    if (!u[k]) return cudaErrorInvalidValue;
    p.p[k] = u[k];
  }
  spherical_kh_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, centers, p);
  return cudaGetLastError();
}

}  // namespace

extern "C" {

int t8b200_cartesian_brick_connectivity(t8b200_cart_conn* out, int is_f64, int dim, int level, int bx, int by, int bz,
                                        int nranks, int rank, void* stream) {
  if (!out || (dim != 2 && dim != 3) || level < 0 || dim * level > 40 || nranks < 1 || rank < 0 || rank >= nranks ||
      bx < 1 || by < 1 || bz < 1 || (dim == 2 && bz != 1) || bx > 1024 || by > 1024 || bz > 1024)
    return cudaErrorInvalidValue;
  *out = t8b200_cart_conn{};
  Geo g;
  g.dim = dim; g.level = level; g.nranks = nranks; g.rank = rank;
  g.b[0] = bx; g.b[1] = by; g.b[2] = bz;
  g.N  = ((int64_t)bx * by * bz) << (dim * level);
  g.lo = part_off(g.N, nranks, rank);
  g.hi = part_off(g.N, nranks, rank + 1);
  if (g.hi - g.lo > 0x7fffffff / 8) return cudaErrorInvalidValue;  // t8_locidx_t is 32 bit
  int rc = is_f64 ? build<double>(out, g, (cudaStream_t)stream) : build<float>(out, g, (cudaStream_t)stream);
  if (rc != 0) t8b200_cartesian_connectivity_free(out);
  return rc;
}

int t8b200_cartesian_uniform_connectivity(t8b200_cart_conn* out, int is_f64, int dim, int level, int nranks, int rank,
                                          void* stream) {
  return t8b200_cartesian_brick_connectivity(out, is_f64, dim, level, 1, 1, 1, nranks, rank, stream);
}

void t8b200_cartesian_connectivity_free(t8b200_cart_conn* c) {
  if (!c) return;
  cudaFree(c->ranks); cudaFree(c->indices); cudaFree(c->face_neighbors); cudaFree(c->face_normals);
  cudaFree(c->face_surfaces); cudaFree(c->volumes); cudaFree(c->centroids); cudaFree(c->x_face_neighbors);
  cudaFree(c->x_face_normals); cudaFree(c->x_face_surfaces);
  *c = t8b200_cart_conn{};
}

int t8b200_init_spherical_kelvin_helmholtz_f32(int64_t n, const float* centers, float* const* u, void* stream) {
  return spherical_kh_impl<float>(n, centers, u, stream);
}
int t8b200_init_spherical_kelvin_helmholtz_f64(int64_t n, const double* centers, double* const* u, void* stream) {
  return spherical_kh_impl<double>(n, centers, u, stream);
}
int t8b200_init_kelvin_helmholtz_f32(int dim, int64_t n, const float* centers, float* const* u, void* stream) {
  return kh_impl<float>(dim, n, centers, u, stream);
}
int t8b200_init_kelvin_helmholtz_f64(int dim, int64_t n, const double* centers, double* const* u, void* stream) {
  return kh_impl<double>(dim, n, centers, u, stream);
}
}
