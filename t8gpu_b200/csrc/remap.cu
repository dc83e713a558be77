// Device-side remap of the variables and volumes after t8code adapt / partition (which stay on the host), for sm_100a.
// C ABI in include/t8gpu_b200.h (section 5).  Pure bandwidth kernels: one thread per NEW element / cell, every value
// written straight into the new allocation (the reference stages through a temporary buffer and five device-to-device
// set_variable copies, mesh_manager.inl:290-324, subgrid_mesh_manager.inl:520-552).
//
// Reference behaviour replaced (not translated):
//   t8gpu/mesh/mesh_manager.inl:164-193           adapt_variables_and_volume
//   t8gpu/mesh/mesh_manager.inl:625-643           partition_data
//   t8gpu/mesh/subgrid_mesh_manager.inl:245-425   adapt_volume, adapt_variables (3-D and 2-D)
//   t8gpu/mesh/subgrid_mesh_manager.inl:1216-1283 partition_variable_data, partition_volume_data
#include "../../include/t8gpu_b200.h"
#include "common.cuh"
#include "peer_sync.cuh"
#include "tile_plan.cuh"

namespace {

constexpr int MAXV = 8;
template <typename T>
struct Ptrs { T* p[MAXV]; };
template <typename T>
struct PtrsC { const T* p[MAXV]; };
template <typename T>
struct TablesC { const T* const* p[MAXV]; };
template <typename T>
struct Tables { T* const* p[MAXV]; };

// adapt_data[i] = first old element behind new element i (n_new + 1 entries, mesh_manager.inl:258-281):
//   diff = adapt_data[i+1] - adapt_data[i] : 0 -> i is a child of a refined element (not the last one),
//   1 -> copy OR the last child of a refined element (then adapt_data[i-1] == adapt_data[i]), > 1 -> coarsened family.
template <typename T>
__global__ void __launch_bounds__(256)
adapt_elements_kernel(int nvar, int64_t n_new, const int32_t* __restrict__ ad, PtrsC<T> uo, Ptrs<T> un,
                      const T* __restrict__ vo, T* __restrict__ vn) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_new) return;
  const int a = ad[i], diff = ad[i + 1] - a;
  const int nsum = diff > 1 ? diff : 1;
  // volume factors are the reference's 3-D constants whatever the mesh dimension (mesh_manager.inl:180-183)
  T v = vo[a] * (diff == 0 ? T(0.125) : (diff == 1 ? T(1.0) : T(8.0)));
  if (i > 0 && ad[i - 1] == a) v = vo[a] * T(0.125);
  vn[i] = v;
#pragma unroll
  for (int k = 0; k < MAXV; k++) {
    if (k < nvar) {
      T s = T(0);
      for (int j = 0; j < nsum; j++) s += uo.p[k][a + j] / T(nsum);   // same order and rounding as the reference
      un.p[k][i] = s;
    }
  }
}

// Subgrid: one thread per new cell; DIM = 3: Subgrid<4,4,4>, DIM = 2: Subgrid<4,4>.
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
adapt_cells_kernel(int nvar, int64_t n_new, const int32_t* __restrict__ ad, PtrsC<T> uo, Ptrs<T> un,
                   const T* __restrict__ vo, T* __restrict__ vn) {
  constexpr int S = DIM == 3 ? 64 : 16, NC = 1 << DIM;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e = g / S;
  if (e >= n_new) return;
  const int c = (int)(g % S), i = c & 3, j = (c >> 2) & 3, k = DIM == 3 ? c >> 4 : 0;
  const int a = ad[e], diff = ad[e + 1] - a;
  const bool child = diff == 0 || (e > 0 && ad[e - 1] == a);
  if (c == 0) {   // adapt_volume, subgrid_mesh_manager.inl:245-290
    const T fr = DIM == 3 ? T(0.125) : T(0.25), fc = DIM == 3 ? T(8.0) : T(4.0);
    T v = vo[a] * (diff == 0 ? fr : (diff == 1 ? T(1.0) : fc));
    if (e > 0 && ad[e - 1] == a) v = vo[a] * fr;
    vn[e] = v;
  }
  if (child) {   // injection from the parent's cell (i/2, j/2, k/2) of octant (I,J,K), :310-330
    int ri = 0;
    while (e - ri >= 0 && ad[e - ri] == a) ri++;
    const int I = (ri - 1) & 1, J = ((ri - 1) >> 1) & 1, K = ((ri - 1) >> 2) & 1;
    const int64_t src = (int64_t)a * S + (I * 2 + i / 2) + 4 * (J * 2 + j / 2) + (DIM == 3 ? 16 * (K * 2 + k / 2) : 0);
#pragma unroll
    for (int l = 0; l < MAXV; l++)
      if (l < nvar) un.p[l][g] = uo.p[l][src];
  } else if (diff > 1) {   // mean of the 2^DIM fine cells, summed in the reference's order (ii, jj, kk), :332-352
    const int I = i >> 1, J = j >> 1, K = k >> 1, z = I | (J << 1) | (K << 2);
    const int64_t b = (int64_t)(a + z) * S;
#pragma unroll
    for (int l = 0; l < MAXV; l++) {
      if (l < nvar) {
        T s = T(0);
        for (int ii = 0; ii < 2; ii++)
          for (int jj = 0; jj < 2; jj++)
            for (int kk = 0; kk < (DIM == 3 ? 2 : 1); kk++)
              s += uo.p[l][b + (2 * (i & 1) + ii) + 4 * (2 * (j & 1) + jj) + (DIM == 3 ? 16 * (2 * (k & 1) + kk) : 0)];
        un.p[l][g] = s / T(NC);
      }
    }
  } else {
    const int64_t src = (int64_t)a * S + c;
#pragma unroll
    for (int l = 0; l < MAXV; l++)
      if (l < nvar) un.p[l][g] = uo.p[l][src];
  }
}

// new element / cell i <- old (ranks[e], indices[e]) through the [var][rank] pointer tables (possibly peer GPUs)
template <typename T>
__global__ void __launch_bounds__(256)
partition_kernel(int nvar, int64_t n_new_cells, int cpe, const int32_t* __restrict__ ranks,
                 const int32_t* __restrict__ indices, Ptrs<T> un, TablesC<T> uo, T* __restrict__ vn,
                 const T* const* __restrict__ vo) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_new_cells) return;
  const int64_t e = g / cpe;
  const int     c = (int)(g % cpe), rk = ranks[e];
  const int64_t src = (int64_t)indices[e] * cpe + c;
#pragma unroll
  for (int l = 0; l < MAXV; l++)
    if (l < nvar) un.p[l][g] = uo.p[l][rk][src];
  if (c == 0 && vn) vn[e] = vo[rk][indices[e]];
}

// Ghost tail of a plan built with t8b200_plan_create_ghost_tail: tail entry j of every variable row <- entry idx[j] of
// rank rk[j]'s row (peer memory over NVLink).  One thread per (entry, variable): every load of the exchange is in flight
// at once, so the pull is bound by NVLink bandwidth, not by the latency a stage kernel would expose per chunk.
template <typename T>
__global__ void __launch_bounds__(256)
ghost_pull_kernel(int nvar, int64_t n_pull, int64_t tail, const int32_t* __restrict__ rk, const int32_t* __restrict__ idx,
                  Ptrs<T> own, TablesC<T> all) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_pull * nvar) return;
  const int64_t j = g % n_pull;
  const int     l = (int)(g / n_pull);
#pragma unroll
  for (int k = 0; k < MAXV; k++)
    if (k == l) own.p[k][tail + j] = all.p[k][rk[j]][idx[j]];
}

// The same exchange as a PUSH from the owner: entry e of the send list, rows_all[k][dst_rank[e]][dst_idx[e]] =
// rows_own[k][src_idx[e]].  The list is sorted by destination, so the remote stores of a warp are consecutive (whole
// 128-byte lines over NVLink, posted writes) and only the local gather is scattered -- the pull reads isolated 8-byte
// values out of 32-byte sectors of the peer's memory and waits for each of them.
template <typename T>
__global__ void __launch_bounds__(256)
ghost_push_kernel(int nvar, int64_t n_send, const int32_t* __restrict__ src, const int32_t* __restrict__ drk,
                  const int32_t* __restrict__ dix, PtrsC<T> own, Tables<T> all) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_send * nvar) return;
  const int64_t e = g % n_send;
  const int     l = (int)(g / n_send);
#pragma unroll
  for (int k = 0; k < MAXV; k++)
    if (k == l) all.p[k][drk[e]][dix[e]] = own.p[k][src[e]];
}

// The push and the barrier that publishes it in ONE launch: every CTA pushes its entries, releases them at system
// scope and counts itself; the last CTA to finish runs the barrier protocol of t8b200_peer_barrier (signal every
// peer, wait for every peer).  One launch per stage instead of two between consecutive stage kernels.
template <typename T>
__global__ void __launch_bounds__(256)
ghost_push_barrier_kernel(int nvar, int64_t n_send, const int32_t* __restrict__ src, const int32_t* __restrict__ drk,
                          const int32_t* __restrict__ dix, PtrsC<T> own, Tables<T> all, unsigned* counter, int nranks,
                          int rank, long long epoch, t8b200::PeerSlot* const* mailboxes, const void* value,
                          int value_is_f64, void* out_max) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_send * nvar) {
    const int64_t e = g % n_send;
    const int     l = (int)(g / n_send);
#pragma unroll
    for (int k = 0; k < MAXV; k++)
      if (k == l) all.p[k][drk[e]][dix[e]] = own.p[k][src[e]];
  }
  __shared__ int last;
  __syncthreads();
  if (threadIdx.x == 0) {
    // one system-scope fence per CTA: the CTA barrier orders the other threads' remote stores before it (cumulativity),
    // so they are visible before the CTA counts itself
    __threadfence_system();
    last = atomicAdd(counter, 1u) + 1u == gridDim.x;
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < 32) {
    __threadfence_system();   // acquire the counts of the other CTAs, release in front of the epoch stores
    t8b200::peer_barrier_warp(threadIdx.x, nranks, rank, epoch, mailboxes, value, value_is_f64, out_max);
    if (threadIdx.x == 0) *counter = 0u;   // the next launch on this stream starts from zero
  }
}

template <typename T>
int adapt_impl(int dim_subgrid, int nvar, int64_t n_new, const int32_t* ad, const T* const* uo, T* const* un,
               const T* vo, T* vn, void* stream) {
  if (nvar < 1 || nvar > MAXV || n_new < 0 || (dim_subgrid != 0 && dim_subgrid != 2 && dim_subgrid != 3))
    return cudaErrorInvalidValue;
  if (n_new == 0) return 0;
  if (!ad || !uo || !un || !vo || !vn) return cudaErrorInvalidValue;
  PtrsC<T> o{};
  Ptrs<T>  n{};
  for (int k = 0; k < nvar; k++) { o.p[k] = uo[k]; n.p[k] = un[k]; }
  cudaStream_t st = (cudaStream_t)stream;
  if (dim_subgrid == 0) {
    adapt_elements_kernel<T><<<(unsigned)((n_new + 255) / 256), 256, 0, st>>>(nvar, n_new, ad, o, n, vo, vn);
  } else if (dim_subgrid == 3) {
    adapt_cells_kernel<T, 3><<<(unsigned)((n_new * 64 + 255) / 256), 256, 0, st>>>(nvar, n_new, ad, o, n, vo, vn);
  } else {
    adapt_cells_kernel<T, 2><<<(unsigned)((n_new * 16 + 255) / 256), 256, 0, st>>>(nvar, n_new, ad, o, n, vo, vn);
  }
  return cudaGetLastError();
}

template <typename T>
int partition_impl(int nvar, int64_t n_new, int cpe, const int32_t* ranks, const int32_t* indices, T* const* un,
                   const T* const* const* uo_all, T* vn, const T* const* vo_all, void* stream) {
  if (nvar < 1 || nvar > MAXV || n_new < 0 || cpe < 1) return cudaErrorInvalidValue;
  if (n_new == 0) return 0;
  if (!ranks || !indices || !un || !uo_all || (vn && !vo_all)) return cudaErrorInvalidValue;
  Ptrs<T>    n{};
  TablesC<T> o{};
  for (int k = 0; k < nvar; k++) { n.p[k] = un[k]; o.p[k] = uo_all[k]; }
  const int64_t cells = n_new * cpe;
  partition_kernel<T><<<(unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nvar, cells, cpe, ranks,
                                                                                         indices, n, o, vn, vo_all);
  return cudaGetLastError();
}
}  // namespace

template <typename T>
static int ghost_pull_impl(const t8b200_plan* P, int nvar, T* const* rows, const T* const* const* rows_all, void* stream) {
  if (!P || !P->ghost_tail || P->host_only || nvar < 1 || nvar > MAXV || (P->is_f64 != (sizeof(T) == 8))) return cudaErrorInvalidValue;
  if (P->n_pull == 0) return 0;
  if (!rows || !rows_all) return cudaErrorInvalidValue;
  Ptrs<T>    o{};
  TablesC<T> a{};
  for (int k = 0; k < nvar; k++) { o.p[k] = rows[k]; a.p[k] = rows_all[k]; }
  const int64_t total = P->n_pull * nvar;
  ghost_pull_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nvar, P->n_pull, P->n_local,
                                                                                          P->pull_rank, P->pull_idx, o, a);
  return cudaGetLastError();
}

template <typename T>
static int ghost_push_impl(int nvar, int64_t n_send, const int32_t* src, const int32_t* drk, const int32_t* dix,
                           const T* const* rows, T* const* const* rows_all, void* stream) {
  if (nvar < 1 || nvar > MAXV || n_send < 0) return cudaErrorInvalidValue;
  if (n_send == 0) return 0;
  if (!src || !drk || !dix || !rows || !rows_all) return cudaErrorInvalidValue;
  PtrsC<T>  o{};
  Tables<T> a{};
  for (int k = 0; k < nvar; k++) { o.p[k] = rows[k]; a.p[k] = rows_all[k]; }
  const int64_t total = n_send * nvar;
  ghost_push_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nvar, n_send, src, drk, dix, o, a);
  return cudaGetLastError();
}

template <typename T>
static int ghost_push_barrier_impl(int nvar, int64_t n_send, const int32_t* src, const int32_t* drk, const int32_t* dix,
                                   const T* const* rows, T* const* const* rows_all, unsigned* counter, int nranks, int rank,
                                   long long epoch, void* const* mailboxes, const void* value, void* out_max, void* stream) {
  if (nvar < 1 || nvar > MAXV || n_send < 0 || nranks < 1 || nranks > 32 || rank < 0 || rank >= nranks || epoch <= 0)
    return cudaErrorInvalidValue;
  if (!counter || !mailboxes || !rows || !rows_all || (n_send > 0 && (!src || !drk || !dix))) return cudaErrorInvalidValue;
  PtrsC<T>  o{};
  Tables<T> a{};
  for (int k = 0; k < nvar; k++) { o.p[k] = rows[k]; a.p[k] = rows_all[k]; }
  const int64_t total = std::max<int64_t>(n_send * nvar, 1);   // a rank with nothing to send still takes part in the barrier
  ghost_push_barrier_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      nvar, n_send, src, drk, dix, o, a, counter, nranks, rank, epoch, (t8b200::PeerSlot* const*)mailboxes, value,
      sizeof(T) == 8, out_max);
  return cudaGetLastError();
}

extern "C" {
int t8b200_ghost_push_f32(int nvar, int64_t n_send, const int32_t* src_idx, const int32_t* dst_rank,
                          const int32_t* dst_idx, const float* const* rows, float* const* const* rows_all, void* stream) {
  return ghost_push_impl<float>(nvar, n_send, src_idx, dst_rank, dst_idx, rows, rows_all, stream);
}
int t8b200_ghost_push_f64(int nvar, int64_t n_send, const int32_t* src_idx, const int32_t* dst_rank,
                          const int32_t* dst_idx, const double* const* rows, double* const* const* rows_all,
                          void* stream) {
  return ghost_push_impl<double>(nvar, n_send, src_idx, dst_rank, dst_idx, rows, rows_all, stream);
}
int t8b200_ghost_push_barrier_f32(int nvar, int64_t n_send, const int32_t* src_idx, const int32_t* dst_rank,
                                  const int32_t* dst_idx, const float* const* rows, float* const* const* rows_all,
                                  unsigned* counter_dev, int nranks, int rank, long long epoch, void* const* mailboxes_dev,
                                  const float* value_dev, float* out_max_dev, void* stream) {
  return ghost_push_barrier_impl<float>(nvar, n_send, src_idx, dst_rank, dst_idx, rows, rows_all, counter_dev, nranks, rank,
                                        epoch, mailboxes_dev, value_dev, out_max_dev, stream);
}
int t8b200_ghost_push_barrier_f64(int nvar, int64_t n_send, const int32_t* src_idx, const int32_t* dst_rank,
                                  const int32_t* dst_idx, const double* const* rows, double* const* const* rows_all,
                                  unsigned* counter_dev, int nranks, int rank, long long epoch, void* const* mailboxes_dev,
                                  const double* value_dev, double* out_max_dev, void* stream) {
  return ghost_push_barrier_impl<double>(nvar, n_send, src_idx, dst_rank, dst_idx, rows, rows_all, counter_dev, nranks,
                                         rank, epoch, mailboxes_dev, value_dev, out_max_dev, stream);
}
int t8b200_ghost_pull_f32(const t8b200_plan* plan, int nvar, float* const* rows, const float* const* const* rows_all,
                          void* stream) {
  return ghost_pull_impl<float>(plan, nvar, rows, rows_all, stream);
}
int t8b200_ghost_pull_f64(const t8b200_plan* plan, int nvar, double* const* rows, const double* const* const* rows_all,
                          void* stream) {
  return ghost_pull_impl<double>(plan, nvar, rows, rows_all, stream);
}
int t8b200_adapt_remap_f32(int subgrid_dim, int nvar, int64_t n_new, const int32_t* adapt_data,
                           const float* const* vars_old, float* const* vars_new, const float* vol_old, float* vol_new,
                           void* stream) {
  return adapt_impl<float>(subgrid_dim, nvar, n_new, adapt_data, vars_old, vars_new, vol_old, vol_new, stream);
}
int t8b200_adapt_remap_f64(int subgrid_dim, int nvar, int64_t n_new, const int32_t* adapt_data,
                           const double* const* vars_old, double* const* vars_new, const double* vol_old,
                           double* vol_new, void* stream) {
  return adapt_impl<double>(subgrid_dim, nvar, n_new, adapt_data, vars_old, vars_new, vol_old, vol_new, stream);
}
int t8b200_partition_remap_f32(int nvar, int64_t n_new, int cells_per_element, const int32_t* ranks,
                               const int32_t* indices, float* const* vars_new, const float* const* const* vars_old_all,
                               float* vol_new, const float* const* vol_old_all, void* stream) {
  return partition_impl<float>(nvar, n_new, cells_per_element, ranks, indices, vars_new, vars_old_all, vol_new,
                               vol_old_all, stream);
}
int t8b200_partition_remap_f64(int nvar, int64_t n_new, int cells_per_element, const int32_t* ranks,
                               const int32_t* indices, double* const* vars_new,
                               const double* const* const* vars_old_all, double* vol_new,
                               const double* const* vol_old_all, void* stream) {
  return partition_impl<double>(nvar, n_new, cells_per_element, ranks, indices, vars_new, vars_old_all, vol_new,
                                vol_old_all, stream);
}
}
