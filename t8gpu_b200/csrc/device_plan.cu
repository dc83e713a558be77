// Tile plan built ON THE DEVICE from device-resident connectivity arrays (SURVEY f-2), for the meshes whose chunks are
// all structured (box_layout.cuh): 256 consecutive elements forming an 8 x 8 x 4 box of same-size hexahedra with 256
// single same-size face neighbours -- every uniform forest / brick partition (BASELINE configs 2, and the weak-scaling
// meshes), i.e. the plans the stage kernel of structured.cu runs without any other plan array.  No device -> host copy
// of the connectivity, no host loop: three kernels and a sort of the ghost keys.  Anything else (hanging faces, walls,
// general normals, ragged ends) is reported as cudaErrorNotSupported and goes through the host builder of
// tile_plan.cuh, whose arrays 13-15 (s_rec, s_halo, s_hrank) and 17-18 (ghost tail) this builder reproduces bit for
// bit (tests/test_device_plan_gpu.py).
//
// Reference behaviour replaced: the part of MeshManager::compute_connectivity_information that uploads the face arrays
// for the kernels (t8gpu/mesh/mesh_manager.inl:442-480) -- here the arrays are already on the device (csrc/cartesian.cu
// builds them there) and are re-laid out into the per-chunk halo lists in place.
#include <thrust/copy.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/scan.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/t8gpu_b200.h"
#include "box_layout.cuh"
#include "common.cuh"
#include "mesh_faces.cuh"
#include "plan_block.cuh"
#include "subgrid_faces.cuh"
#include "tile_plan.cuh"

using namespace t8b200;

namespace {

enum Flag { NOT_AXIS = 0, MANY_AREAS, SIDE_TAKEN, NOT_BOX, N_BOUNDARY, N_FLAGS };

template <typename T>
struct DevConn {
  int64_t        n_local;
  int32_t        nf, nx;
  const int32_t* nbr;
  const T *      normals, *areas;
  const int32_t* xnbr;
  const T *      xnormals, *xareas;
};

// face f -> the two (element, side) entries of the element -> neighbour table; a side that is written twice (2:1
// hanging faces seen from the coarse element) or a normal / an area that does not fit disqualifies the mesh
template <typename T>
__global__ void __launch_bounds__(256) neighbour_table_kernel(DevConn<T> c, T area0, int32_t* __restrict__ nb6, int* flags) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= (int64_t)c.nf + c.nx) return;
  int32_t  l, r;
  const T* n;
  T        a;
  if (f < c.nf) { l = c.nbr[2 * f]; r = c.nbr[2 * f + 1]; n = c.normals + 3 * f; a = c.areas[f]; }
  else { const int64_t g = f - c.nf; l = c.xnbr[2 * g]; r = c.xnbr[2 * g + 1]; n = c.xnormals + 3 * g; a = c.xareas[g]; }
  int code = -1;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const T o1 = n[(d + 1) % 3], o2 = n[(d + 2) % 3];
    if (o1 == T(0) && o2 == T(0) && (n[d] == T(1) || n[d] == T(-1))) code = 2 * d + (n[d] > T(0) ? 1 : 0);
  }
  if (code < 0) { flags[NOT_AXIS] = 1; return; }
  if (a != area0) flags[MANY_AREAS] = 1;
  const int axis = code >> 1, plus = code & 1;   // the normal points l -> r: r sits on l's (+axis if plus) side
  if (l < c.n_local && atomicCAS(nb6 + (int64_t)l * 6 + 2 * axis + plus, -1, r) != -1) flags[SIDE_TAKEN] = 1;
  if (r < c.n_local && atomicCAS(nb6 + (int64_t)r * 6 + 2 * axis + 1 - plus, -1, l) != -1) flags[SIDE_TAKEN] = 1;
}

// CTA = block of 256 consecutive elements: is it the box of layout L, and if so its halo list in thread order
template <class L>
__global__ void __launch_bounds__(256)
box_kernel(int64_t n_local, const int32_t* __restrict__ nb6, const int32_t* __restrict__ ranks,
           const int32_t* __restrict__ indices, int multi, int my_rank, const int16_t* __restrict__ thread_of_slot,
           int32_t* __restrict__ s_rec, int32_t* __restrict__ s_halo, int32_t* __restrict__ s_hrank, int* flags) {
  __shared__ int32_t outside[256];
  const int     b = blockIdx.x, t = threadIdx.x;
  const int64_t e0 = (int64_t)b * 256, e = e0 + t;
  bool ok = true, bnd = false;
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const int32_t n = nb6[e * 6 + 2 * d + side];
      const bool    edge = side == 0 ? L::at_lower(t, d) : L::at_upper(t, d);
      if (n < 0) { ok = false; continue; }
      if (!edge) {
        ok = ok && n == (int32_t)e0 + (side == 0 ? L::lower_own(t, d) : L::upper_own(t, d));
      } else {
        ok = ok && (n < e0 || n >= e0 + 256);
        const int h = thread_of_slot[L::halo_slot(d, side, L::compact(t, d))];
        outside[h] = n;
        int32_t rk = my_rank, ix = n;
        if (multi) { rk = ranks[n]; ix = indices[n]; }
        s_halo[(int64_t)b * 256 + h] = ix;
        if (multi) s_hrank[(int64_t)b * 256 + h] = rk;
        bnd = bnd || rk != my_rank;
      }
    }
  // the 256 outside neighbours must be 256 different elements (a coarser neighbour would serve four faces): bitonic
  // sort, then compare neighbours
  __syncthreads();
  for (int k = 2; k <= 256; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int p = t ^ j;
      if (p > t) {
        const int32_t x = outside[t], y = outside[p];
        if (((t & k) == 0) == (x > y)) { outside[t] = y; outside[p] = x; }
      }
      __syncthreads();
    }
  if (t > 0 && outside[t] == outside[t - 1]) ok = false;
  const int all_ok = __syncthreads_and(ok ? 1 : 0), any_bnd = __syncthreads_or(bnd ? 1 : 0);
  if (t == 0) {
    if (!all_ok) flags[NOT_BOX] = 1;
    if (any_bnd) atomicAdd(flags + N_BOUNDARY, 1);
    s_rec[4 * b] = (int32_t)e0; s_rec[4 * b + 1] = 0; s_rec[4 * b + 2] = b; s_rec[4 * b + 3] = any_bnd;
  }
}


// Subgrid<4,4,4>: CTA = 4 consecutive elements = 256 cells.  The chunk is structured when the elements are the 2 x 2 x 1
// arrangement of the layout (el = ex + 2 ey: Morton siblings with the same z) and every element behind the box boundary
// is a single same-level one; halo entries are CELLS of those elements (index in the owner's cell rows).
__global__ void __launch_bounds__(256)
subgrid_box_kernel(int64_t n_elem, const int32_t* __restrict__ nb6, const int32_t* __restrict__ ranks,
                   const int32_t* __restrict__ indices, int multi, int my_rank, const int16_t* __restrict__ thread_of_slot,
                   int32_t* __restrict__ s_rec, int32_t* __restrict__ s_halo, int32_t* __restrict__ s_hrank, int* flags) {
  using L = SubgridBox;
  __shared__ long long outside[256];
  const int     b = blockIdx.x, t = threadIdx.x;
  const int64_t e0 = (int64_t)b * 4;
  const int     el = t >> 6, i = t & 3, j = (t >> 2) & 3, k = (t >> 4) & 3;
  const int64_t e = e0 + el;
  bool ok = true, bnd = false;
  // the four elements sit as the layout says: +x of element 0 / 2 is element 1 / 3, +y of element 0 / 1 is 2 / 3
  if (t < 4) {
    const int ex = t & 1, ey = t >> 1;
    const int32_t xn = nb6[(e0 + t) * 6 + (ex ? 0 : 1)], yn = nb6[(e0 + t) * 6 + 2 + (ey ? 0 : 1)];
    ok = xn == (int32_t)(e0 + (t ^ 1)) && yn == (int32_t)(e0 + (t ^ 2));
  }
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const bool edge = side == 0 ? L::at_lower(t, d) : L::at_upper(t, d);
      if (!edge) continue;
      const int32_t n = nb6[e * 6 + 2 * d + side];
      if (n < 0 || (n >= e0 && n < e0 + 4)) { ok = false; continue; }
      // the cell across the face: same tangential coordinates, the other end of the element along d
      const int c = side == 0 ? 3 : 0;
      const int cell = d == 0 ? c + 4 * j + 16 * k : d == 1 ? i + 4 * c + 16 * k : i + 4 * j + 16 * c;
      int32_t rk = my_rank, ix = n;
      if (multi) { rk = ranks[n]; ix = indices[n]; }
      const int h = thread_of_slot[L::halo_slot(d, side, L::compact(t, d))];
      outside[h] = ((long long)rk << 40) | ((long long)ix * 64 + cell);
      s_halo[(int64_t)b * 256 + h] = ix * 64 + cell;
      if (multi) s_hrank[(int64_t)b * 256 + h] = rk;
      bnd = bnd || rk != my_rank;
    }
  __syncthreads();
  for (int kk = 2; kk <= 256; kk <<= 1)
    for (int jj = kk >> 1; jj > 0; jj >>= 1) {
      const int p = t ^ jj;
      if (p > t) {
        const long long x = outside[t], y = outside[p];
        if (((t & kk) == 0) == (x > y)) { outside[t] = y; outside[p] = x; }
      }
      __syncthreads();
    }
  if (t > 0 && outside[t] == outside[t - 1]) ok = false;
  const int all_ok = __syncthreads_and(ok ? 1 : 0), any_bnd = __syncthreads_or(bnd ? 1 : 0);
  if (t == 0) {
    if (!all_ok) flags[NOT_BOX] = 1;
    if (any_bnd) atomicAdd(flags + N_BOUNDARY, 1);
    s_rec[4 * b] = b * 256; s_rec[4 * b + 1] = 0; s_rec[4 * b + 2] = b; s_rec[4 * b + 3] = any_bnd;
  }
  (void)n_elem;
}

// every element of the same size, every face between elements of the same level
template <typename T>
__global__ void uniform_check_kernel(int64_t n_elem, const T* __restrict__ vol, int64_t nfl, const int32_t* __restrict__ ld,
                                     int64_t nxl, const int32_t* __restrict__ xld, int* flags) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n_elem && vol[g] != vol[0]) flags[MANY_AREAS] = 1;
  if (g < nfl && ld[g] != 0) flags[NOT_BOX] = 1;
  if (g < nxl && xld[g] != 0) flags[NOT_BOX] = 1;
}

__global__ void ghost_keys_kernel(int64_t n, const int32_t* __restrict__ halo, const int32_t* __restrict__ hrank, int me,
                                  unsigned long long* __restrict__ keys, unsigned long long* count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || halo[i] < 0 || hrank[i] == me) return;
  keys[atomicAdd(count, 1ull)] = ((unsigned long long)(uint32_t)hrank[i] << 32) | (uint32_t)halo[i];
}
__global__ void redirect_kernel(int64_t n, int32_t* __restrict__ halo, int32_t* __restrict__ hrank, int me,
                                const unsigned long long* __restrict__ keys, int64_t nkeys, int64_t n_local) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || halo[i] < 0 || hrank[i] == me) return;
  const unsigned long long k = ((unsigned long long)(uint32_t)hrank[i] << 32) | (uint32_t)halo[i];
  int64_t lo = 0, hi = nkeys;   // lower bound
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
  halo[i]  = (int32_t)(n_local + lo);
  hrank[i] = me;
}
struct IsBoundary {
  const int32_t* s_rec;
  __device__ bool operator()(int q) const { return s_rec[4 * (size_t)q + 3] != 0; }
};
__global__ void split_keys_kernel(int64_t n, const unsigned long long* __restrict__ keys, int32_t* __restrict__ rk,
                                  int32_t* __restrict__ ix) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rk[i] = (int32_t)(keys[i] >> 32);
  ix[i] = (int32_t)(keys[i] & 0xFFFFFFFFull);
}

template <typename P>
struct DevFree {   // frees scratch allocations on every exit path
  P* p = nullptr;
  ~DevFree() { cudaFree(p); }
};

// cells = false: MeshManager elements (256 elements per chunk, MortonBox);  cells = true: Subgrid<4,4,4> (4 elements =
// 256 cells per chunk, SubgridBox; vol / ld / xld: device per-element volumes and level differences of the faces)
template <typename T>
int device_plan_impl(t8b200_plan** out, int flags_in, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb,
                     const int32_t* nbr, const T* normals, const T* areas, const int32_t* ranks, const int32_t* indices,
                     int32_t nx, const int32_t* xnbr, const T* xnormals, const T* xareas, void* stream,
                     bool cells = false, const T* vol = nullptr, const int32_t* ld = nullptr,
                     const int32_t* xld = nullptr) {
  if (!out || n_local < 0 || n_ghost < 0 || nf < 0 || nb < 0 || nx < 0) return cudaErrorInvalidValue;
  if ((nf > 0 && (!nbr || !normals || !areas)) || (nx > 0 && (!xnbr || !xnormals || !xareas))) return cudaErrorInvalidValue;
  if (n_ghost > 0 && (!ranks || !indices)) return cudaErrorInvalidValue;
  // cheap disqualifiers first: walls, a ragged last block, nothing to do
  const int64_t per_chunk = cells ? 4 : 256, S = cells ? 64 : 1;
  if (nb != 0 || n_local == 0 || (n_local % per_chunk) != 0 || nf == 0 || (n_local + n_ghost) * S > 0x7FFFFF00LL) return cudaErrorNotSupported;
  if (cells && (!vol || !ld || (nx > 0 && !xld))) return cudaErrorInvalidValue;
  cudaStream_t  st      = (cudaStream_t)stream;
  const int     nchunks = (int)(n_local / per_chunk);
  const bool    multi   = n_ghost > 0;
  DevFree<int32_t> nb6;
  DevFree<int>     flags;
  DevFree<int16_t> inv;
  T8B_TRY(cudaMalloc(&nb6.p, sizeof(int32_t) * 6 * (size_t)n_local));
  T8B_TRY(cudaMalloc(&flags.p, sizeof(int) * N_FLAGS));
  T8B_TRY(cudaMemsetAsync(nb6.p, 0xFF, sizeof(int32_t) * 6 * (size_t)n_local, st));
  T8B_TRY(cudaMemsetAsync(flags.p, 0, sizeof(int) * N_FLAGS, st));
  int16_t inv_h[SubgridBox::NSLOT];
  for (int i = 0; i < SubgridBox::NSLOT; i++) inv_h[i] = -1;
  for (int h = 0; h < 256; h++) inv_h[cells ? SubgridBox::thread_slot(h) : MortonBox::thread_slot(h)] = (int16_t)h;
  T8B_TRY(cudaMalloc(&inv.p, sizeof(inv_h)));
  T8B_TRY(cudaMemcpyAsync(inv.p, inv_h, sizeof(inv_h), cudaMemcpyHostToDevice, st));
  T   area0 = T(0), vol0 = T(0);
  int me    = 0;
  T8B_TRY(cudaMemcpyAsync(&area0, areas, sizeof(T), cudaMemcpyDeviceToHost, st));
  if (cells) T8B_TRY(cudaMemcpyAsync(&vol0, vol, sizeof(T), cudaMemcpyDeviceToHost, st));
  if (multi) T8B_TRY(cudaMemcpyAsync(&me, ranks, sizeof(int), cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaStreamSynchronize(st));

  t8b200_plan* P = new t8b200_plan();
  struct Guard { t8b200_plan* p; ~Guard() { if (p) t8b200_plan_destroy(p); } } guard{P};
  P->is_f64 = sizeof(T) == 8; P->n_local = n_local * S; P->n_chunks = nchunks; P->multi = multi ? 1 : 0; P->my_rank = me;
  P->ghost_tail = (flags_in >> 1) & 1;
  P->cmp = 1; P->n_areas = 1; P->box_layout = cells ? 1 : 0; P->n_struct = nchunks; P->n_generic = 0; P->s_area0 = 0;
  T cell_area = area0;
  if (cells) {
    // faces between cells: (cbrt(vol) / 4)^2 inside an element, face_surface / 16 across elements (kernels.inl:352-354,
    // :786-787) -- one value on a uniform forest, or this builder does not apply
    P->vol_shift = 6; P->vol_scale = 1.0 / 64.0;
    // (libm's cbrt is off by an ulp on exact cubes: the exact root when there is one, as the host builder does)
    T c = std::cbrt(vol0);
    for (T t : {std::nextafter(c, T(0)), std::nextafter(c, T(2) * c)})
      if (t * t * t == vol0) c = t;
    const T inner = (c / T(4)) * (c / T(4));
    cell_area     = area0 / T(16);
    if (inner != cell_area) return cudaErrorNotSupported;
  }
  P->max_halo = 256; P->max_faces = BoxCommon::NFLUX; P->hs = 256; P->fs = BoxCommon::NFLUX; P->ms = MS; P->mf = MF;
  P->n_halo = (int64_t)nchunks * 256; P->n_records = (int64_t)nchunks * BoxCommon::NFLUX;
  P->smem_bytes = sizeof(T) * ((size_t)NCELLQ * (cells ? SubgridBox::NSLOT : MortonBox::NSLOT) + 5 * (size_t)BoxCommon::NFLUX);
  auto dev_alloc = [&](auto** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes + 32);
    if (e == cudaSuccess) P->dev_bytes += (int64_t)bytes + 32;
    return e;
  };
  T8B_TRY(dev_alloc(&P->s_rec, sizeof(int32_t) * 4 * (size_t)nchunks));
  T8B_TRY(dev_alloc(&P->s_halo, sizeof(int32_t) * 256 * (size_t)nchunks));
  if (multi) T8B_TRY(dev_alloc(&P->s_hrank, sizeof(int32_t) * 256 * (size_t)nchunks));
  T8B_TRY(dev_alloc((T**)&P->area_tab, sizeof(T)));
  T8B_TRY(cudaMemcpyAsync(P->area_tab, &cell_area, sizeof(T), cudaMemcpyHostToDevice, st));

  DevConn<T> c{n_local, nf, nx, nbr, normals, areas, xnbr, xnormals, xareas};
  const int64_t ntot = (int64_t)nf + nx;
  neighbour_table_kernel<T><<<(unsigned)((ntot + 255) / 256), 256, 0, st>>>(c, area0, nb6.p, flags.p);
  if (cells) {
    const int64_t m = std::max<int64_t>(n_local, std::max<int64_t>(nf, nx));
    uniform_check_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, st>>>(n_local, vol, nf, ld, nx, xld, flags.p);
    subgrid_box_kernel<<<nchunks, 256, 0, st>>>(n_local, nb6.p, ranks, indices, multi ? 1 : 0, me, inv.p, P->s_rec,
                                                P->s_halo, P->s_hrank, flags.p);
  } else {
    box_kernel<MortonBox><<<nchunks, 256, 0, st>>>(n_local, nb6.p, ranks, indices, multi ? 1 : 0, me, inv.p, P->s_rec,
                                                   P->s_halo, P->s_hrank, flags.p);
  }
  T8B_TRY(cudaGetLastError());
  int fl[N_FLAGS];
  T8B_TRY(cudaMemcpyAsync(fl, flags.p, sizeof(fl), cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaStreamSynchronize(st));
  if (fl[NOT_AXIS] || fl[MANY_AREAS] || fl[SIDE_TAKEN] || fl[NOT_BOX]) return cudaErrorNotSupported;
  if (multi) {
    P->nb_struct = fl[N_BOUNDARY];
    if (P->nb_struct == 0) {   // no ghosts read at all: one nominal boundary chunk (every rank must signal)
      const int32_t one = 1;
      T8B_TRY(cudaMemcpyAsync(P->s_rec + 3, &one, sizeof(one), cudaMemcpyHostToDevice, st));
      P->nb_struct = 1;
    }
    // ids of the partition-boundary chunks, ascending (the "boundary pass" of the split stage launches)
    T8B_TRY(dev_alloc(&P->blist, sizeof(int32_t) * (size_t)P->nb_struct));
    thrust::copy_if(thrust::cuda::par.on(st), thrust::counting_iterator<int>(0), thrust::counting_iterator<int>(nchunks),
                    thrust::device_ptr<int32_t>(P->blist), IsBoundary{P->s_rec});
    T8B_TRY(cudaGetLastError());
    T8B_TRY(cudaStreamSynchronize(st));
  }
  if (P->ghost_tail && multi) {
    // distinct (owner rank, remote index) pairs of the ghost entries, sorted -> tail slots; entries redirected
    const int64_t nh = (int64_t)nchunks * 256;
    DevFree<unsigned long long> keys, count;
    T8B_TRY(cudaMalloc(&keys.p, sizeof(unsigned long long) * (size_t)nh));
    T8B_TRY(cudaMalloc(&count.p, sizeof(unsigned long long)));
    T8B_TRY(cudaMemsetAsync(count.p, 0, sizeof(unsigned long long), st));
    ghost_keys_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(nh, P->s_halo, P->s_hrank, me, keys.p, count.p);
    unsigned long long nk = 0;
    T8B_TRY(cudaMemcpyAsync(&nk, count.p, sizeof(nk), cudaMemcpyDeviceToHost, st));
    T8B_TRY(cudaStreamSynchronize(st));
    thrust::device_ptr<unsigned long long> kb(keys.p);
    thrust::sort(thrust::cuda::par.on(st), kb, kb + nk);
    const int64_t nu = thrust::unique(thrust::cuda::par.on(st), kb, kb + nk) - kb;
    if (nu + P->n_local > 0x7FFFFF00LL) return cudaErrorInvalidValue;
    if (nu > 0) {
      redirect_kernel<<<(unsigned)((nh + 255) / 256), 256, 0, st>>>(nh, P->s_halo, P->s_hrank, me, keys.p, nu, P->n_local);
      T8B_TRY(dev_alloc(&P->pull_rank, sizeof(int32_t) * (size_t)nu));
      T8B_TRY(dev_alloc(&P->pull_idx, sizeof(int32_t) * (size_t)nu));
      split_keys_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, st>>>(nu, keys.p, P->pull_rank, P->pull_idx);
    }
    P->n_pull = nu;
    T8B_TRY(cudaGetLastError());
    T8B_TRY(cudaStreamSynchronize(st));
  }
  guard.p = nullptr;
  *out    = P;
  return 0;
}


// =====================================================================================================================
// Any other mesh (hanging faces, walls, general normals, ragged ends, blocks that must be split): the generic builder.
// One CUDA thread runs the per-block program of plan_block.cuh for one block of 256 elements -- the programs are
// independent, branchy integer work over ~1000 faces each, latency-bound; thousands of them in flight hide it -- between
// data-parallel passes over the faces (classification, bucketing) and over the chunks (launch lists, ghost tail).
// Device -> host traffic: counters, the <= 256 distinct areas and two flag bytes per chunk; never the connectivity.
// Every array equals the host builder's (tests/test_device_plan_gpu.py).
// =====================================================================================================================
enum GFlag { G_NOT_AXIS = 0, G_MANY_AREAS, G_BAD_FACE, G_PROGRAM_ERROR, G_MY_RANK, G_N };
constexpr unsigned long long AREA_EMPTY = ~0ull;
constexpr int                AREA_SLOTS = 257;

template <typename T>
__device__ inline unsigned long long area_bits(T a) {
  if constexpr (sizeof(T) == 8) return (unsigned long long)__double_as_longlong((double)a);
  else return (unsigned long long)__float_as_uint((float)a);
}

// per face: normal class and area into the set of distinct areas (lock-free insert), faces per block
template <typename T, class Src>
__global__ void __launch_bounds__(256) classify_kernel(Src src, int64_t ntot, int64_t n_local, int multi,
                                                       unsigned long long* area_set, int* flags,
                                                       unsigned long long* blk_cnt, unsigned long long* keys) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f == 0 && n_local > 0 && multi) { int32_t rk = 0, ix = 0; src.owner(0, rk, ix); flags[G_MY_RANK] = rk; }
  if (f >= ntot) return;
  T nrm[3], a;
  src.geometry(f, nrm, a);
  if (pb::axis_code_hd(nrm) < 0) flags[G_NOT_AXIS] = 1;
  else {
    const unsigned long long bits = area_bits(a);
    int i = 0;
    for (; i < AREA_SLOTS; i++) {
      unsigned long long v = *(volatile unsigned long long*)(area_set + i);
      if (v == bits) break;
      if (v == AREA_EMPTY) {
        v = atomicCAS(area_set + i, AREA_EMPTY, bits);
        if (v == AREA_EMPTY || v == bits) break;
      }
    }
    if (i == AREA_SLOTS) flags[G_MANY_AREAS] = 1;
  }
  int32_t l, r;
  src.endpoints(f, l, r);
  const int64_t cl = l < n_local ? l / pb::EC : -1, cr = (r >= 0 && r < n_local) ? r / pb::EC : -1;
  if (keys) {   // block << 40 | face: sorted, the candidates of every block in ascending face order; ~0 = none
    keys[2 * f]     = cl >= 0 ? ((unsigned long long)cl << 40) | (unsigned long long)f : ~0ull;
    keys[2 * f + 1] = (cr >= 0 && cr != cl) ? ((unsigned long long)cr << 40) | (unsigned long long)f : ~0ull;
  }
  if (cl < 0 && cr < 0) { flags[G_BAD_FACE] = 1; return; }
  if (cl >= 0) atomicAdd(blk_cnt + cl, 1ull);
  if (cr >= 0 && cr != cl) atomicAdd(blk_cnt + cr, 1ull);
}

__global__ void __launch_bounds__(256) mask_keys_kernel(int64_t n, unsigned long long* keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && keys[i] != ~0ull) keys[i] &= (1ull << 40) - 1;
}

template <typename T, class Src, bool FILL>
__global__ void __launch_bounds__(64) block_pass_kernel(Src src, pb::Params<T> pr, unsigned char* arena, int64_t nprog,
                                                        int64_t blk0, int64_t nblocks, const unsigned long long* face_off,
                                                        int64_t* rec, pb::Counts* cn, const unsigned long long* bases,
                                                        pb::Out<T> out, int* flags) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, blk = blk0 + t;
  if (t >= nprog || blk >= nblocks) return;
  const pb::Ws w = pb::Ws::carve(arena, nprog, t, true);
  pb::Counts   c;
  pb::block_program<T, Src, FILL>(src, pr, w, blk, rec + face_off[blk], (int64_t)(face_off[blk + 1] - face_off[blk]), c,
                                  FILL ? (int64_t)bases[blk] : 0, FILL ? (int64_t)bases[(nblocks + 1) + blk] : 0,
                                  FILL ? (int64_t)bases[2 * (nblocks + 1) + blk] : 0, out);
  if (c.rc || (FILL && c.chunks != cn[blk].chunks)) flags[G_PROGRAM_ERROR] = 1;
  if (!FILL) cn[blk] = c;
}

// ---------------------------------------------------------------------------------------------------------------------
// The block program of plan_block.cuh executed by one WARP with its workspace in shared memory (the default; the
// thread-per-block kernel above stays selectable with T8B200_DEVICE_PLAN=serial and is the form the host emulation
// runs).  Step by step the same program, each loop in its data-parallel form:
//   faces of the range        ballot + prefix compaction over the block's (sorted) candidate list
//   distinct outside elements atomicCAS inserts into the open-addressing table; sorted by counting ranks
//   records in kernel order   bitonic sort of the (group, left slot, right slot, position) keys
//   element -> face table     bitonic sort of the (slot, entry) pairs; rank inside the slot's run = entry position
//   structured test           atomicCAS fills of the lower / upper neighbour tables, then one lane per element
// ---------------------------------------------------------------------------------------------------------------------
namespace wp {
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr unsigned long long FACE_MASK = (1ull << 40) - 1;   // candidate key = block << 40 | face id

struct Smem {
  unsigned long long key[pb::MF];
  int32_t            ends[2 * pb::MF];   // endpoints; then the (slot, entry) pairs; then the neighbour tables
  int32_t            pos[pb::MF];
  int32_t            ht_key[pb::HT];
  uint16_t           ht_val[pb::HT];
  int32_t            halo[pb::MS - pb::EC], halo_sorted[pb::MS - pb::EC];
  int32_t            el_cnt[pb::EC], el_off[pb::EC], ovf_loc[pb::EC];
  int                nh, seg[4], flag;
};

template <typename U>
__device__ inline void warp_bitonic(U* a, int n, int lane) {   // n: power of two >= 32
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += 32) {
        const int p = i ^ j;
        if (p > i) {
          const U x = a[i], y = a[p];
          if (((i & k) == 0) == (x > y)) { a[i] = y; a[p] = x; }
        }
      }
      __syncwarp();
    }
}

template <typename T, class Src, bool FILL>
__device__ int chunk_warp(const Src& src, const pb::Params<T>& pr, Smem& sm, const unsigned long long* cand, int64_t ncand,
                          bool whole, int64_t b0, int64_t b1, pb::Counts& cn, int64_t c, int64_t oo_at, int64_t oe_at,
                          const pb::Out<T>& out) {
  constexpr int EC = pb::EC, MS = pb::MS, MF = pb::MF, HT = pb::HT, ELL = pb::ELL, MAXH = pb::MS - pb::EC;
  const int lane = threadIdx.x & 31;
  // the faces of the range, ascending in the face id
  int nfc = 0;
  for (int64_t q0 = 0; q0 < ncand; q0 += 32) {
    const int64_t q = q0 + lane;
    bool          keep = false;
    int32_t       l = 0, r = 0;
    if (q < ncand) {
      src.endpoints((int64_t)(cand[q] & FACE_MASK), l, r);
      keep = whole || (l >= b0 && l < b1) || (r >= b0 && r < b1);
    }
    const unsigned m  = __ballot_sync(FULL, keep);
    const int      at = nfc + __popc(m & ((1u << lane) - 1u));
    if (keep && at < MF) { sm.pos[at] = (int32_t)q; sm.ends[2 * at] = l; sm.ends[2 * at + 1] = r; }
    nfc += __popc(m);
    if (nfc > pr.max_faces_allowed) return 1;
  }
  // distinct elements outside the range
  for (int i = lane; i < HT; i += 32) sm.ht_key[i] = -1;
  if (lane == 0) sm.nh = 0;
  __syncwarp();
  for (int j = lane; j < 2 * nfc; j += 32) {
    const int32_t id = sm.ends[j];
    if (id < 0 || (id >= b0 && id < b1)) continue;
    unsigned k = pb::halo_hash(id);
    while (*(volatile int*)&sm.nh <= MAXH) {   // more than MAXH distinct: the chunk does not fit, stop filling the table
      const int32_t old = atomicCAS(&sm.ht_key[k], -1, id);
      if (old == -1) {
        const int s = atomicAdd(&sm.nh, 1);
        if (s < MAXH) sm.halo[s] = id;
        break;
      }
      if (old == id) break;
      k = (k + 1) & (HT - 1);
    }
  }
  __syncwarp();
  const int nh = sm.nh;
  if (nh > pr.max_halo_allowed) return 1;
  // entries per own slot, their exclusive prefix sums (all entries / entries beyond ELL)
  for (int i = lane; i < EC; i += 32) sm.el_cnt[i] = 0;
  __syncwarp();
  for (int j = lane; j < 2 * nfc; j += 32) {
    const int32_t id = sm.ends[j];
    if (id >= b0 && id < b1) atomicAdd(&sm.el_cnt[id - b0], 1);
  }
  __syncwarp();
  int s1 = 0, s2 = 0;
  for (int i = 0; i < 8; i++) { const int v = sm.el_cnt[lane * 8 + i]; s1 += v; s2 += v > ELL ? v - ELL : 0; }
  int p1 = s1, p2 = s2;
  for (int d = 1; d < 32; d <<= 1) {
    const int t1 = __shfl_up_sync(FULL, p1, d), t2 = __shfl_up_sync(FULL, p2, d);
    if (lane >= d) { p1 += t1; p2 += t2; }
  }
  const int n_pairs = __shfl_sync(FULL, p1, 31), n_ovf = __shfl_sync(FULL, p2, 31);
  {
    int e1 = p1 - s1, e2 = p2 - s2;
    for (int i = 0; i < 8; i++) {
      const int v = sm.el_cnt[lane * 8 + i];
      sm.el_off[lane * 8 + i] = e1; sm.ovf_loc[lane * 8 + i] = e2;
      e1 += v; e2 += v > ELL ? v - ELL : 0;
    }
  }
  __syncwarp();
  if (!FILL) {
    cn.chunks++;
    cn.max_halo  = nh > cn.max_halo ? nh : cn.max_halo;
    cn.max_faces = nfc > cn.max_faces ? nfc : cn.max_faces;
    cn.sum_halo += nh;
    cn.sum_faces += nfc;
    if (n_ovf) { cn.ovf_off += EC + 1; cn.ovf_ent += n_ovf; }
    return 0;
  }

  int32_t* H = out.hdr + 8 * c;
  // halo: sorted by counting ranks (ids are distinct), owners, slot of every id
  for (int i = lane; i < nh; i += 32) {
    const int32_t id = sm.halo[i];
    int           rk = 0;
    for (int j = 0; j < nh; j++) rk += sm.halo[j] < id;
    sm.halo_sorted[rk] = id;
  }
  __syncwarp();
  int bad = 0;
  for (int h = lane; h < nh; h += 32) {
    const int32_t id = sm.halo_sorted[h];
    int32_t       rk = 0, ix = id;
    if (pr.multi) src.owner(id, rk, ix);
    else if (id >= pr.n_local) bad = 1;
    out.halo_elem[c * out.HS + h] = ix;
    if (pr.multi) out.halo_rank[c * out.HS + h] = rk;
    unsigned k = pb::halo_hash(id);
    while (sm.ht_key[k] != id) k = (k + 1) & (HT - 1);
    sm.ht_val[k] = (uint16_t)(EC + h);
  }
  if (__any_sync(FULL, bad)) return -1;
  if (lane < 4) sm.seg[lane] = 0;
  __syncwarp();
  auto slot_of = [&](int32_t id) -> int {
    if (id >= b0 && id < b1) return (int)(id - b0);
    unsigned k = pb::halo_hash(id);
    while (sm.ht_key[k] != id) k = (k + 1) & (HT - 1);
    return sm.ht_val[k];
  };
  // records in kernel order
  int n2 = 32;
  while (n2 < nfc) n2 <<= 1;
  for (int j = lane; j < n2; j += 32) {
    if (j >= nfc) { sm.key[j] = ~0ull; continue; }
    const int32_t l = sm.ends[2 * j], r = sm.ends[2 * j + 1];
    int           sl = slot_of(l), sr = r < 0 ? 0xFFFF : slot_of(r), grp = 0;
    if (pr.cmp) {
      T nrm[3], a;
      src.geometry((int64_t)(cand[sm.pos[j]] & FACE_MASK), nrm, a);
      const int code = pb::axis_code_hd(nrm);
      grp = r < 0 ? 3 : code >> 1;
      if (r < 0) sr = 0xFFF8 | code;
      else if (!(code & 1)) { const int t = sl; sl = sr; sr = t; }
      atomicAdd(&sm.seg[grp], 1);
    }
    sm.key[j] = ((unsigned long long)(grp * MS + sl) << 32) | ((unsigned long long)sr << 16) | (unsigned long long)j;
  }
  __syncwarp();
  warp_bitonic(sm.key, n2, lane);
  if (lane == 0) {
    H[0] = (int32_t)b0;
    H[1] = (int32_t)(b1 - b0);
    H[2] = nh | (nfc << 16);
    H[3] = sm.seg[0] | ((sm.seg[0] + sm.seg[1]) << 16);
    H[4] = sm.seg[0] + sm.seg[1] + sm.seg[2];
    H[5] = n_ovf ? (int32_t)oo_at : -1;
    H[6] = n_ovf ? (int32_t)oe_at : 0;
  }
  if (n_ovf) {
    for (int i = lane; i < EC; i += 32) out.ovf_off[oo_at + i] = (uint16_t)sm.ovf_loc[i];
    if (lane == 0) out.ovf_off[oo_at + EC] = (uint16_t)n_ovf;
  }
  // emission: records, geometry, (slot, entry) pairs
  uint32_t* pairs = reinterpret_cast<uint32_t*>(sm.ends);
  int       n2p   = 64;
  while (n2p < 2 * nfc) n2p <<= 1;
  for (int i = 2 * nfc + lane; i < n2p; i += 32) pairs[i] = 0xFFFFFFFFu;
  auto area_index = [&](T a) {
    int lo = 0, hi = pr.n_areas - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (pr.area_tab[mid] < a) lo = mid + 1; else hi = mid; }
    return lo;
  };
  int area0 = -1;
  if (pr.cmp && nfc > 0) {
    T nrm[3], a;
    src.geometry((int64_t)(cand[sm.pos[(int)(sm.key[0] & 0xFFFFu)]] & FACE_MASK), nrm, a);
    area0 = area_index(a);
  }
  int mismatch = 0;
  for (int jj = lane; jj < nfc; jj += 32) {
    const unsigned long long kk = sm.key[jj];
    const int j = (int)(kk & 0xFFFFu), sr = (int)((kk >> 16) & 0xFFFFu), gs = (int)(kk >> 32), grp = gs / MS, sl = gs % MS;
    T         nrm[3], a;
    src.geometry((int64_t)(cand[sm.pos[j]] & FACE_MASK), nrm, a);
    if (pr.cmp) {
      const int ai = area_index(a);
      out.face_ai[c * out.FS + jj] = (uint8_t)ai;
      mismatch |= ai != area0;
    } else {
      out.fnx[c * out.FS + jj] = nrm[0]; out.fny[c * out.FS + jj] = nrm[1]; out.fnz[c * out.FS + jj] = nrm[2];
      out.far[c * out.FS + jj] = a;
    }
    const uint32_t axis_bits = (pr.cmp && grp < 3) ? (uint32_t)grp << 14 : 0u;
    out.face_lr[c * out.FS + jj] = (uint32_t)sl | axis_bits | ((uint32_t)sr << 16);
    pairs[2 * jj]     = sl < EC ? ((uint32_t)sl << 16) | (uint32_t)(jj << 1) : 0xFFFFFFFFu;
    pairs[2 * jj + 1] = sr < EC ? ((uint32_t)sr << 16) | (uint32_t)((jj << 1) | 1) : 0xFFFFFFFFu;
  }
  const bool uniform = pr.cmp && !__any_sync(FULL, mismatch);
  __syncwarp();
  warp_bitonic(pairs, n2p, lane);
  for (int i = lane; i < n_pairs; i += 32) {
    const uint32_t p = pairs[i];
    const int      slot = (int)(p >> 16), k = i - sm.el_off[slot];
    const uint16_t en = (uint16_t)(p & 0xFFFFu);
    if (k < ELL) out.ell[(b0 + slot) * ELL + k] = en;
    else out.ovf_ent[oe_at + sm.ovf_loc[slot] + (k - ELL)] = en;
  }
  if (lane == 0) H[7] = (uniform && area0 >= 0) ? area0 : -1;
  __syncwarp();
  // structured?
  int structured = 0;
  if (pr.box_layout >= 0 && pr.cmp && uniform && area0 >= 0 && b1 - b0 == 256 && (b0 & 255) == 0 && sm.seg[3] == 0 &&
      nh == 256 && nfc == BoxCommon::NFLUX) {
    int32_t *lo = sm.ends, *hi = sm.ends + 3 * EC;
    for (int i = lane; i < 6 * EC; i += 32) sm.ends[i] = -1;
    if (lane == 0) sm.flag = 1;
    __syncwarp();
    for (int jj = lane; jj < nfc; jj += 32) {
      const unsigned long long kk = sm.key[jj];
      const int sr = (int)((kk >> 16) & 0xFFFFu), gs = (int)(kk >> 32), d = gs / MS, sl = gs % MS;
      if (sl < 256 && atomicCAS(&hi[d * EC + sl], -1, sr) != -1) sm.flag = 0;
      if (sr < 256 && atomicCAS(&lo[d * EC + sr], -1, sl) != -1) sm.flag = 0;
    }
    __syncwarp();
    auto test = [&](auto tag) {
      using L = decltype(tag);
      for (int t = lane; t < 256; t += 32)
        for (int d = 0; d < 3; d++) {
          const int l = lo[d * EC + t], u = hi[d * EC + t];
          if ((L::at_lower(t, d) ? l < 256 : l != L::lower_own(t, d)) || (L::at_upper(t, d) ? u < 256 : u != L::upper_own(t, d))) {
            sm.flag = 0;
            continue;
          }
          if (L::at_lower(t, d)) {
            const int h = pr.thread_of_slot[L::halo_slot(d, 0, L::compact(t, d))];
            out.s_halo[c * 256 + h] = out.halo_elem[c * out.HS + (l - 256)];
            if (pr.multi) out.s_hrank[c * 256 + h] = out.halo_rank[c * out.HS + (l - 256)];
          }
          if (L::at_upper(t, d)) {
            const int h = pr.thread_of_slot[L::halo_slot(d, 1, L::compact(t, d))];
            out.s_halo[c * 256 + h] = out.halo_elem[c * out.HS + (u - 256)];
            if (pr.multi) out.s_hrank[c * 256 + h] = out.halo_rank[c * out.HS + (u - 256)];
          }
        }
    };
    if (sm.flag) { if (pr.box_layout == 1) test(SubgridBox{}); else test(MortonBox{}); }
    __syncwarp();
    structured = sm.flag;
  }
  if (lane == 0) out.s_flag[c] = (uint8_t)structured;
  __syncwarp();
  cn.chunks++;
  if (n_ovf) { cn.ovf_off += EC + 1; cn.ovf_ent += n_ovf; }
  return 0;
}
}  // namespace wp

// CTA = one warp; walks the blocks blockIdx.x, blockIdx.x + gridDim.x, ...
template <typename T, class Src, bool FILL>
__global__ void __launch_bounds__(32) block_warp_kernel(Src src, pb::Params<T> pr, int64_t nblocks,
                                                        const unsigned long long* face_off, const unsigned long long* keys,
                                                        pb::Counts* cn, const unsigned long long* bases, pb::Out<T> out,
                                                        int* flags) {
  __shared__ wp::Smem sm;
  struct Range { int64_t b0, b1; };
  for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    Range         todo[40];
    int           sp = 0;
    const int64_t blk_b0 = blk * pb::EC, blk_b1 = blk * pb::EC + pb::EC < pr.n_local ? blk * pb::EC + pb::EC : pr.n_local;
    const unsigned long long* cand  = keys + face_off[blk];
    const int64_t             ncand = (int64_t)(face_off[blk + 1] - face_off[blk]);
    const int64_t cb = FILL ? (int64_t)bases[blk] : 0, ob = FILL ? (int64_t)bases[(nblocks + 1) + blk] : 0,
                  eb = FILL ? (int64_t)bases[2 * (nblocks + 1) + blk] : 0;
    pb::Counts c{0, 0, 0, 0, 0, 0, 0, 0};
    todo[sp++] = {blk_b0, blk_b1};
    while (sp > 0) {
      const Range rg = todo[--sp];
      const int   rc = wp::chunk_warp<T, Src, FILL>(src, pr, sm, cand, ncand, rg.b1 - rg.b0 == blk_b1 - blk_b0, rg.b0, rg.b1, c,
                                                    cb + c.chunks, ob + c.ovf_off, eb + c.ovf_ent, out);
      __syncwarp();
      if (rc == 1 && rg.b1 - rg.b0 > 1) {
        const int64_t mid = (rg.b0 + rg.b1) / 2;
        todo[sp].b0 = mid;   todo[sp++].b1 = rg.b1;
        todo[sp].b0 = rg.b0; todo[sp++].b1 = mid;
        continue;
      }
      if (rc) { c.rc = 1; break; }
    }
    if (threadIdx.x == 0) {
      if (c.rc || (FILL && c.chunks != cn[blk].chunks)) flags[G_PROGRAM_ERROR] = 1;
      if (!FILL) cn[blk] = c;
    }
    __syncwarp();
  }
}

// counts of the blocks -> three arrays for the scans (chunks, overflow offsets, overflow entries) + totals
__global__ void __launch_bounds__(256) counts_kernel(int64_t nblocks, const pb::Counts* cn, unsigned long long* bases,
                                                     unsigned long long* totals) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const pb::Counts c = cn[b];
  bases[b] = c.chunks; bases[(nblocks + 1) + b] = c.ovf_off; bases[2 * (nblocks + 1) + b] = c.ovf_ent;
  atomicMax(totals + 0, (unsigned long long)c.max_halo);
  atomicMax(totals + 1, (unsigned long long)c.max_faces);
  atomicAdd(totals + 2, (unsigned long long)c.sum_halo);
  atomicAdd(totals + 3, (unsigned long long)c.sum_faces);
  if (c.chunks > 1) totals[4] = 1;
}

// CTA = chunk: does it read another rank's elements?  (structured chunks: their 256-entry list, else the halo list)
__global__ void __launch_bounds__(256) chunk_flags_kernel(int HS, const int32_t* __restrict__ halo_elem,
                                                          const int32_t* __restrict__ halo_rank,
                                                          const uint8_t* __restrict__ s_flag, int me, uint8_t* bflag) {
  const int64_t c = blockIdx.x;
  bool          b = false;
  for (int h = threadIdx.x; h < HS; h += 256) b = b || (halo_elem[c * HS + h] >= 0 && halo_rank[c * HS + h] != me);
  const int any = __syncthreads_or(b ? 1 : 0);
  if (threadIdx.x == 0) bflag[c] = (uint8_t)any;
  (void)s_flag;
}

// CTA = structured chunk at launch position q: record + its 256-entry halo list from the per-chunk scratch
__global__ void __launch_bounds__(256) gather_structured_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ hdr,
                                                                const int32_t* __restrict__ s_halo_tmp,
                                                                const int32_t* __restrict__ s_hrank_tmp, int32_t* s_rec,
                                                                int32_t* s_halo, int32_t* s_hrank) {
  const int64_t q = blockIdx.x;
  const int32_t o = order[q], c = o & 0x3FFFFFFF;
  s_halo[q * 256 + threadIdx.x] = s_halo_tmp[(int64_t)c * 256 + threadIdx.x];
  if (s_hrank) s_hrank[q * 256 + threadIdx.x] = s_hrank_tmp[(int64_t)c * 256 + threadIdx.x];
  if (threadIdx.x == 0) {
    s_rec[4 * q] = hdr[8 * (int64_t)c]; s_rec[4 * q + 1] = hdr[8 * (int64_t)c + 7]; s_rec[4 * q + 2] = c;
    s_rec[4 * q + 3] = (o >> 30) & 1;
  }
}

// One allocation carved into 256-byte aligned pieces (cudaMalloc / cudaFree cost far more than the kernels of a plan
// build: measured 6-16 ms for the 14 plan arrays of a 0.7 M-element forest).  A Pool without base measures.
struct Pool {
  unsigned char* base = nullptr;
  size_t         used = 0;
  template <typename U>
  U* take(size_t count) {
    const size_t off = used;
    used += (count * sizeof(U) + 32 + 255) & ~(size_t)255;   // 32 bytes of slack: 16-byte granular prefetch hints
    return base ? reinterpret_cast<U*>(base + off) : nullptr;
  }
};
template <typename U>
struct Piece { U* p = nullptr; };   // a piece of a Pool (same `.p` access as DevFree, no ownership)

template <typename T, class Src>
int generic_device_plan(t8b200_plan* P, int64_t n_local, bool multi, const Src& src, cudaStream_t st) {
  static_assert(pb::EC == EC && pb::MS == MS && pb::MF == MF && pb::ELL == ELL, "plan_block.cuh constants");
  const int64_t nblocks = (n_local + EC - 1) / EC, ntot = src.num_faces();
  auto          pol     = thrust::cuda::par.on(st);
  // T8B200_DEVICE_PLAN=serial: one THREAD per block runs pb::block_program (the form the host emulation checks)
  const bool serial = getenv("T8B200_DEVICE_PLAN") && getenv("T8B200_DEVICE_PLAN")[0] == 's';
  static const bool timing = getenv("T8B200_PLAN_TIMING") != nullptr;   // phases to stderr (adds a synchronisation each)
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(st);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[t8b200 device plan] %-34s %.3f ms\n", what, 1e3 * std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  P->n_local = n_local; P->multi = multi ? 1 : 0;
  auto dev_alloc = [&](auto** p, size_t bytes, int fill) {
    cudaError_t e = cudaMalloc(p, bytes + 32);
    if (e != cudaSuccess) return e;
    P->dev_bytes += (int64_t)bytes + 32;
    return cudaMemsetAsync(*p, fill, bytes + 32, st);
  };
  // ---- faces: classes, distinct areas, buckets per block
  if (nblocks >= (1LL << 23) || ntot >= (1LL << 40)) return cudaErrorInvalidValue;
  Piece<unsigned long long> area_set, blk, totals, face_off, keys;   // blk: 3 x (nb+1) bases after the count pass
  Piece<int>                flags;
  Piece<pb::Counts>         cn;
  Piece<int16_t>            inv;
  auto carve_scratch = [&](Pool& pool) {
    area_set.p = pool.take<unsigned long long>(AREA_SLOTS);
    flags.p    = pool.take<int>(G_N);
    blk.p      = pool.take<unsigned long long>(3 * (size_t)(nblocks + 1));
    face_off.p = pool.take<unsigned long long>((size_t)(nblocks + 1));
    totals.p   = pool.take<unsigned long long>(8);
    cn.p       = pool.take<pb::Counts>((size_t)nblocks);
    inv.p      = pool.take<int16_t>(SubgridBox::NSLOT);
    keys.p     = pool.take<unsigned long long>(2 * (size_t)std::max<int64_t>(ntot, 1));
  };
  DevFree<unsigned char> scratch;
  {
    Pool measure;
    carve_scratch(measure);
    T8B_TRY(cudaMalloc(&scratch.p, measure.used));
    Pool pool{scratch.p, 0};
    carve_scratch(pool);
  }
  T8B_TRY(cudaMemsetAsync(area_set.p, 0xFF, 8 * AREA_SLOTS, st));
  T8B_TRY(cudaMemsetAsync(flags.p, 0, sizeof(int) * G_N, st));
  T8B_TRY(cudaMemsetAsync(blk.p, 0, 8 * 3 * (size_t)(nblocks + 1), st));
  T8B_TRY(cudaMemsetAsync(face_off.p, 0, 8 * (size_t)(nblocks + 1), st));
  T8B_TRY(cudaMemsetAsync(totals.p, 0, 8 * 8, st));
  const unsigned fgrid = (unsigned)((ntot + 255) / 256);
  // candidate faces of every block: keys block << 40 | face (a face between two blocks appears in both), sorted
  if (ntot > 0) classify_kernel<T, Src><<<fgrid, 256, 0, st>>>(src, ntot, n_local, multi ? 1 : 0, area_set.p, flags.p, face_off.p, keys.p);
  thrust::device_ptr<unsigned long long> fo(face_off.p), kp(keys.p);
  thrust::exclusive_scan(pol, fo, fo + nblocks + 1, fo);
  thrust::sort(pol, kp, kp + 2 * ntot);
  int64_t* rec = reinterpret_cast<int64_t*>(keys.p);   // serial mode: the face ids alone
  if (serial && ntot > 0) mask_keys_kernel<<<(unsigned)((2 * ntot + 255) / 256), 256, 0, st>>>(2 * ntot, keys.p);
  lap("classify + key sort");
  // ---- COUNT pass: one program per block, in batches that bound the workspace
  const int64_t nprog = std::min<int64_t>(nblocks, 65536);
  const unsigned wgrid = (unsigned)std::min<int64_t>(nblocks, 148 * 32);
  DevFree<unsigned char> arena;
  if (serial) T8B_TRY(cudaMalloc(&arena.p, (size_t)pb::Ws::bytes_per_program * (size_t)nprog + 64));
  int max_halo_allowed = MS - EC;
  if (const char* t = getenv("T8B200_TEST_MAX_HALO")) max_halo_allowed = std::min(max_halo_allowed, std::max(8, atoi(t)));
  const int box_layout = P->vol_shift == 6 ? 1 : P->vol_shift == 0 ? 0 : -1;
  P->box_layout        = box_layout;
  int16_t inv_h[SubgridBox::NSLOT];
  for (int i = 0; i < SubgridBox::NSLOT; i++) inv_h[i] = -1;
  for (int h = 0; h < 256; h++) inv_h[box_layout == 1 ? SubgridBox::thread_slot(h) : MortonBox::thread_slot(h)] = (int16_t)h;
  T8B_TRY(cudaMemcpyAsync(inv.p, inv_h, sizeof(inv_h), cudaMemcpyHostToDevice, st));
  pb::Params<T> pr{n_local, multi ? 1 : 0, 0, 0, box_layout, max_halo_allowed, MF - 1, nullptr, inv.p};
  pb::Out<T>    out{};
  if (serial)
    for (int64_t b0 = 0; b0 < nblocks; b0 += nprog)
      block_pass_kernel<T, Src, false><<<(unsigned)((nprog + 63) / 64), 64, 0, st>>>(src, pr, arena.p, nprog, b0, nblocks,
                                                                                    face_off.p, rec, cn.p, blk.p, out, flags.p);
  else
    block_warp_kernel<T, Src, false><<<wgrid, 32, 0, st>>>(src, pr, nblocks, face_off.p, keys.p, cn.p, blk.p, out, flags.p);
  lap("count pass");
  counts_kernel<<<(unsigned)((nblocks + 255) / 256), 256, 0, st>>>(nblocks, cn.p, blk.p, totals.p);
  thrust::device_ptr<unsigned long long> bp(blk.p);
  for (int k = 0; k < 3; k++) thrust::exclusive_scan(pol, bp + k * (nblocks + 1), bp + (k + 1) * (nblocks + 1), bp + k * (nblocks + 1));
  unsigned long long tot[8], ends[3], aset[AREA_SLOTS];
  int                fl[G_N];
  T8B_TRY(cudaMemcpyAsync(tot, totals.p, sizeof(tot), cudaMemcpyDeviceToHost, st));
  for (int k = 0; k < 3; k++) T8B_TRY(cudaMemcpyAsync(&ends[k], blk.p + k * (nblocks + 1) + nblocks, 8, cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaMemcpyAsync(aset, area_set.p, sizeof(aset), cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaMemcpyAsync(fl, flags.p, sizeof(fl), cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaGetLastError());
  T8B_TRY(cudaStreamSynchronize(st));
  if (fl[G_BAD_FACE] || fl[G_PROGRAM_ERROR]) return cudaErrorInvalidValue;
  lap("scans + counters to the host");
  if (multi) P->my_rank = fl[G_MY_RANK];
  const int64_t nchunks = (int64_t)ends[0], n_oo = (int64_t)ends[1], n_oe = (int64_t)ends[2];
  if (n_local > 0x7FFFFF00LL || nchunks * MF > 0x7FFFFF00LL || n_oo > 0x7FFFFF00LL || n_oe > 0x7FFFFF00LL) return cudaErrorInvalidValue;
  std::vector<T> area_tab;
  bool           cmp = !fl[G_NOT_AXIS] && !fl[G_MANY_AREAS];
  if (cmp) {
    for (int i = 0; i < AREA_SLOTS && aset[i] != AREA_EMPTY; i++) {
      T a;
      if constexpr (sizeof(T) == 8) memcpy(&a, &aset[i], 8); else { const uint32_t u = (uint32_t)aset[i]; memcpy(&a, &u, 4); }
      area_tab.push_back(a);
    }
    if (area_tab.size() > 256) cmp = false;
    std::sort(area_tab.begin(), area_tab.end());
  }
  const int max_halo = (int)tot[0], max_faces = (int)tot[1];
  const int HS = std::max(32, (max_halo + 31) / 32 * 32), FS = std::max(32, (max_faces + 31) / 32 * 32);
  P->n_chunks = (int)nchunks; P->split = tot[4] ? 1 : 0; P->hs = HS; P->fs = FS; P->max_halo = max_halo; P->max_faces = max_faces;
  P->n_halo = (int64_t)tot[2]; P->n_records = (int64_t)tot[3]; P->ms = MS; P->mf = MF;
  P->smem_bytes = sizeof(T) == 8 ? 8 * ((size_t)NCELLQ * MS + 5 * (size_t)MF) : 32 * (size_t)MS + 20 * (size_t)MF;
  P->cmp = cmp ? 1 : 0; P->n_areas = cmp ? (int)area_tab.size() : 0;
  // ---- plan arrays in one allocation (padding: halo -1, element -> face table 0xFFFF, everything else 0) and the
  // per-chunk scratch of the fill pass in another
  size_t ones_end = 0;
  auto carve_plan = [&](Pool& pool) {
    P->halo_elem = pool.take<int32_t>((size_t)nchunks * HS);
    P->ell       = reinterpret_cast<uint4*>(pool.take<uint16_t>((size_t)std::max<int64_t>(n_local, 1) * ELL));
    ones_end     = pool.used;                                     // [0, ones_end): 0xFF, the rest: 0
    P->hdr       = pool.take<int32_t>(8 * (size_t)nchunks);
    if (multi) P->halo_rank = pool.take<int32_t>((size_t)nchunks * HS);
    P->face_lr = pool.take<uint32_t>((size_t)nchunks * FS);
    if (cmp) {
      P->face_ai  = pool.take<uint8_t>((size_t)nchunks * FS);
      P->area_tab = pool.take<T>(std::max<size_t>(area_tab.size(), 1));
    } else {
      P->fnx   = pool.take<T>((size_t)nchunks * FS);
      P->fny   = pool.take<T>((size_t)nchunks * FS);
      P->fnz   = pool.take<T>((size_t)nchunks * FS);
      P->farea = pool.take<T>((size_t)nchunks * FS);
    }
    P->ovf_off = pool.take<uint16_t>((size_t)std::max<int64_t>(n_oo, 1));
    P->ovf_ent = pool.take<uint16_t>((size_t)std::max<int64_t>(n_oe, 1));
  };
  {
    Pool measure;
    carve_plan(measure);
    T8B_TRY(cudaMalloc(&P->pool, measure.used));
    P->dev_bytes += (int64_t)measure.used;
    Pool pool{static_cast<unsigned char*>(P->pool), 0};
    carve_plan(pool);
    T8B_TRY(cudaMemsetAsync(P->pool, 0xFF, ones_end, st));
    T8B_TRY(cudaMemsetAsync(static_cast<unsigned char*>(P->pool) + ones_end, 0, measure.used - ones_end, st));
  }
  if (cmp) T8B_TRY(cudaMemcpyAsync(P->area_tab, area_tab.data(), sizeof(T) * area_tab.size(), cudaMemcpyHostToDevice, st));
  P->n_ovf_off = n_oo; P->n_ovf_ent = n_oe;
  Piece<uint8_t> s_flag, bflag;
  Piece<int32_t> s_halo_tmp, s_hrank_tmp;
  auto carve_chunks = [&](Pool& pool) {
    s_flag.p     = pool.take<uint8_t>((size_t)nchunks + 1);
    bflag.p      = pool.take<uint8_t>((size_t)nchunks + 1);
    s_halo_tmp.p = pool.take<int32_t>(256 * (size_t)nchunks + 1);
    if (multi) s_hrank_tmp.p = pool.take<int32_t>(256 * (size_t)nchunks + 1);
  };
  DevFree<unsigned char> chunk_scratch;
  {
    Pool measure;
    carve_chunks(measure);
    T8B_TRY(cudaMalloc(&chunk_scratch.p, measure.used));
    Pool pool{chunk_scratch.p, 0};
    carve_chunks(pool);
  }
  // ---- FILL pass
  pr.cmp = cmp ? 1 : 0; pr.n_areas = (int)area_tab.size(); pr.area_tab = (const T*)P->area_tab;
  out = pb::Out<T>{HS, FS, P->hdr, P->halo_elem, P->halo_rank, P->face_lr, P->face_ai, (T*)P->fnx, (T*)P->fny, (T*)P->fnz,
                   (T*)P->farea, (uint16_t*)P->ell, P->ovf_off, P->ovf_ent, s_flag.p, s_halo_tmp.p, s_hrank_tmp.p};
  if (serial)
    for (int64_t b0 = 0; b0 < nblocks; b0 += nprog)
      block_pass_kernel<T, Src, true><<<(unsigned)((nprog + 63) / 64), 64, 0, st>>>(src, pr, arena.p, nprog, b0, nblocks,
                                                                                   face_off.p, rec, cn.p, blk.p, out, flags.p);
  else
    block_warp_kernel<T, Src, true><<<wgrid, 32, 0, st>>>(src, pr, nblocks, face_off.p, keys.p, cn.p, blk.p, out, flags.p);
  lap("allocations + fill pass");
  const int me = P->my_rank;
  if (multi && nchunks > 0) chunk_flags_kernel<<<(unsigned)nchunks, 256, 0, st>>>(HS, P->halo_elem, P->halo_rank, s_flag.p, me, bflag.p);
  else T8B_TRY(cudaMemsetAsync(bflag.p, 0, (size_t)nchunks + 1, st));
  // ---- launch lists: two flag bytes per chunk to the host, the order back (tile_plan.cuh: arrange)
  std::vector<uint8_t> sf((size_t)nchunks), bf((size_t)nchunks);
  T8B_TRY(cudaMemcpyAsync(sf.data(), s_flag.p, (size_t)nchunks, cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaMemcpyAsync(bf.data(), bflag.p, (size_t)nchunks, cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaMemcpyAsync(fl, flags.p, sizeof(fl), cudaMemcpyDeviceToHost, st));
  T8B_TRY(cudaGetLastError());
  T8B_TRY(cudaStreamSynchronize(st));
  if (fl[G_PROGRAM_ERROR]) return cudaErrorInvalidValue;
  std::vector<int32_t> s_list, g_list, blist;
  lap("chunk flags to the host");
  for (int64_t c = 0; c < nchunks; c++) (sf[c] ? s_list : g_list).push_back((int32_t)c);
  P->n_struct  = (int)s_list.size();
  P->n_generic = P->n_struct ? (int)g_list.size() : (int)nchunks;
  if (!P->n_struct && !multi) g_list.clear();
  if (multi) {
    const bool keep_order = t8b_boundary_order_mode() == 2 || (P->n_struct == nchunks && !P->split);
    auto arrange = [&](std::vector<int32_t>& list, int& n_bnd) {   // tile_plan.cuh: t8b_boundary_order
      std::vector<uint8_t> flag(list.size());
      n_bnd = 0;
      for (size_t q = 0; q < list.size(); q++) { flag[q] = bf[list[q]]; n_bnd += flag[q]; }
      std::vector<int32_t> order;
      order.reserve(list.size());
      for (size_t q : t8b_boundary_order(flag, keep_order)) order.push_back(list[q] | (flag[q] ? 1 << 30 : 0));
      list.swap(order);
    };
    arrange(s_list, P->nb_struct);
    arrange(g_list, P->nb_generic);
    if (P->nb_struct + P->nb_generic == 0 && nchunks > 0) {   // no ghosts at all: one nominal boundary chunk
      if (P->n_struct) { s_list[0] |= 1 << 30; P->nb_struct = 1; } else { g_list[0] |= 1 << 30; P->nb_generic = 1; }
    }
    for (int q = 0; q < P->n_struct; q++) if (s_list[q] >> 30 & 1) blist.push_back(q);
  }
  if (P->n_struct) {
    DevFree<int32_t> order;
    T8B_TRY(cudaMalloc(&order.p, 4 * s_list.size()));
    T8B_TRY(cudaMemcpyAsync(order.p, s_list.data(), 4 * s_list.size(), cudaMemcpyHostToDevice, st));
    T8B_TRY(dev_alloc(&P->s_rec, 4 * 4 * s_list.size(), 0));
    T8B_TRY(dev_alloc(&P->s_halo, 4 * 256 * s_list.size(), 0));
    if (multi) T8B_TRY(dev_alloc(&P->s_hrank, 4 * 256 * s_list.size(), 0));
    gather_structured_kernel<<<(unsigned)s_list.size(), 256, 0, st>>>(order.p, P->hdr, s_halo_tmp.p, s_hrank_tmp.p, P->s_rec,
                                                                     P->s_halo, P->s_hrank);
    int32_t a0 = 0;
    T8B_TRY(cudaMemcpyAsync(&a0, P->s_rec + 1, 4, cudaMemcpyDeviceToHost, st));
    T8B_TRY(cudaStreamSynchronize(st));
    P->s_area0 = a0;
  }
  if (!g_list.empty()) {
    T8B_TRY(dev_alloc(&P->g_list, 4 * g_list.size(), 0));
    T8B_TRY(cudaMemcpyAsync(P->g_list, g_list.data(), 4 * g_list.size(), cudaMemcpyHostToDevice, st));
  }
  if (!blist.empty()) {
    T8B_TRY(dev_alloc(&P->blist, 4 * blist.size(), 0));
    T8B_TRY(cudaMemcpyAsync(P->blist, blist.data(), 4 * blist.size(), cudaMemcpyHostToDevice, st));
  }
  // ---- ghost tail over the halo lists and the structured lists
  if (P->ghost_tail && multi) {
    const int64_t n1 = nchunks * HS, n2 = (int64_t)P->n_struct * 256;
    DevFree<unsigned long long> keys, count;
    T8B_TRY(cudaMalloc(&keys.p, 8 * (size_t)std::max<int64_t>(n1 + n2, 1)));
    T8B_TRY(cudaMalloc(&count.p, 8));
    T8B_TRY(cudaMemsetAsync(count.p, 0, 8, st));
    if (n1) ghost_keys_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(n1, P->halo_elem, P->halo_rank, me, keys.p, count.p);
    if (n2) ghost_keys_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(n2, P->s_halo, P->s_hrank, me, keys.p, count.p);
    unsigned long long nk = 0;
    T8B_TRY(cudaMemcpyAsync(&nk, count.p, 8, cudaMemcpyDeviceToHost, st));
    T8B_TRY(cudaStreamSynchronize(st));
    thrust::device_ptr<unsigned long long> kb(keys.p);
    thrust::sort(pol, kb, kb + nk);
    const int64_t nu = thrust::unique(pol, kb, kb + nk) - kb;
    if (nu + n_local > 0x7FFFFF00LL) return cudaErrorInvalidValue;
    if (nu > 0) {
      if (n1) redirect_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(n1, P->halo_elem, P->halo_rank, me, keys.p, nu, n_local);
      if (n2) redirect_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(n2, P->s_halo, P->s_hrank, me, keys.p, nu, n_local);
      T8B_TRY(dev_alloc(&P->pull_rank, 4 * (size_t)nu, 0));
      T8B_TRY(dev_alloc(&P->pull_idx, 4 * (size_t)nu, 0));
      split_keys_kernel<<<(unsigned)((nu + 255) / 256), 256, 0, st>>>(nu, keys.p, P->pull_rank, P->pull_idx);
    }
    P->n_pull = nu;
  }
  T8B_TRY(cudaGetLastError());
  T8B_TRY(cudaStreamSynchronize(st));
  lap("launch lists + ghost tail");
  return 0;
}

template <typename T>
int generic_mesh_plan(t8b200_plan** out, int ghost_tail, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb,
                      const int32_t* nbr, const T* normals, const T* areas, const int32_t* ranks, const int32_t* indices,
                      int32_t nx, const int32_t* xnbr, const T* xnormals, const T* xareas, void* stream) {
  if (n_local <= 0 || n_local + n_ghost > 0x7FFFFF00LL) return cudaErrorNotSupported;
  t8b200_plan* P = new t8b200_plan();
  struct Guard { t8b200_plan* p; ~Guard() { if (p) t8b200_plan_destroy(p); } } guard{P};
  P->is_f64 = sizeof(T) == 8; P->ghost_tail = ghost_tail ? 1 : 0;
  MeshFaces<T> src{nf, nb, nx, nbr, normals, areas, n_ghost > 0 ? ranks : nullptr, indices, xnbr, xnormals, xareas};
  const int rc = generic_device_plan<T>(P, n_local, n_ghost > 0, src, (cudaStream_t)stream);
  if (rc) return rc;
  guard.p = nullptr;
  *out    = P;
  return 0;
}


template <typename T, int DIM>
__global__ void __launch_bounds__(256) inner_area_kernel(int64_t n, const T* __restrict__ vol, T* __restrict__ inner) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) inner[e] = subgrid_inner_area<T, DIM>(vol[e]);
}

// cell-level plan of Subgrid<4,4,4> / <4,4> over the cell faces (subgrid_faces.cuh), n_local / n_ghost in ELEMENTS
template <typename T, int DIM>
int generic_subgrid_plan(t8b200_plan** out, int ghost_tail, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb,
                         const int32_t* nbr, const T* normals, const T* areas, const int32_t* ld, const int32_t* off,
                         const T* vol, const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                         const T* xnormals, const T* xareas, const int32_t* xld, const int32_t* xoff, void* stream) {
  constexpr int S = DIM == 3 ? 64 : 16;
  if (n_local <= 0) return cudaErrorNotSupported;
  if ((n_local + n_ghost) * S > 0x7FFFFFF0LL) return cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  DevFree<T>   inner;
  T8B_TRY(cudaMalloc(&inner.p, sizeof(T) * (size_t)n_local));
  inner_area_kernel<T, DIM><<<(unsigned)((n_local + 255) / 256), 256, 0, st>>>(n_local, vol, inner.p);
  t8b200_plan* P = new t8b200_plan();
  struct Guard { t8b200_plan* p; ~Guard() { if (p) t8b200_plan_destroy(p); } } guard{P};
  P->is_f64 = sizeof(T) == 8; P->ghost_tail = ghost_tail ? 1 : 0;
  P->vol_shift = DIM == 3 ? 6 : 4; P->vol_scale = DIM == 3 ? 1.0 / 64.0 : 1.0 / 16.0;
  SubgridFaces<T, DIM> src{n_local, nf, nb, nx, nbr, normals, areas, ld, off, vol, ranks, indices, xnbr, xnormals, xareas,
                           xld, xoff, inner.p};
  const int rc = generic_device_plan<T>(P, n_local * S, n_ghost > 0, src, st);
  if (rc) return rc;
  guard.p = nullptr;
  *out    = P;
  return 0;
}

}  // namespace

extern "C" {
int t8b200_plan_create_device(t8b200_plan** out, int is_f64, int ghost_tail, int64_t n_local, int64_t n_ghost,
                              int32_t nf, int32_t nb, const int32_t* nbr, const void* normals, const void* areas,
                              const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                              const void* xnormals, const void* xareas, void* stream) {
  try {
    const int flags = ghost_tail ? 2 : 0;
    const char* mode_env     = getenv("T8B200_DEVICE_PLAN");   // tests: "generic" / "serial" skip the three-kernel builder
    const bool  generic_only = mode_env && (mode_env[0] == 'g' || mode_env[0] == 's');
    int rc = cudaErrorNotSupported;
    if (!generic_only)
      rc = is_f64 ? device_plan_impl<double>(out, flags, n_local, n_ghost, nf, nb, nbr, (const double*)normals,
                                             (const double*)areas, ranks, indices, nx, xnbr, (const double*)xnormals,
                                             (const double*)xareas, stream)
                  : device_plan_impl<float>(out, flags, n_local, n_ghost, nf, nb, nbr, (const float*)normals,
                                            (const float*)areas, ranks, indices, nx, xnbr, (const float*)xnormals,
                                            (const float*)xareas, stream);
    if (rc != cudaErrorNotSupported) return rc;
    // not a structured-only mesh: the generic builder (one program per block of 256 elements)
    if (!out || n_local < 0 || n_ghost < 0 || nf < 0 || nb < 0 || nx < 0) return cudaErrorInvalidValue;
    if ((nf + nb > 0 && (!nbr || !normals || !areas)) || (nx > 0 && (!xnbr || !xnormals || !xareas))) return cudaErrorInvalidValue;
    if (n_ghost > 0 && (!ranks || !indices)) return cudaErrorInvalidValue;
    return is_f64 ? generic_mesh_plan<double>(out, ghost_tail, n_local, n_ghost, nf, nb, nbr, (const double*)normals,
                                              (const double*)areas, ranks, indices, nx, xnbr, (const double*)xnormals,
                                              (const double*)xareas, stream)
                  : generic_mesh_plan<float>(out, ghost_tail, n_local, n_ghost, nf, nb, nbr, (const float*)normals,
                                             (const float*)areas, ranks, indices, nx, xnbr, (const float*)xnormals,
                                             (const float*)xareas, stream);
  } catch (const std::bad_alloc&) {
    return cudaErrorMemoryAllocation;
  } catch (...) {   // thrust reports failed temporary allocations / launches by exception
    cudaGetLastError();
    return cudaErrorUnknown;
  }
}

int t8b200_subgrid_plan_create_device(t8b200_subgrid_plan** out, int is_f64, int dim, int ghost_tail, int64_t n_local,
                                      int64_t n_ghost, int32_t nf, int32_t nb, const int32_t* nbr, const void* normals,
                                      const void* areas, const int32_t* level_diff, const int32_t* offsets,
                                      const void* volumes, const int32_t* ranks, const int32_t* indices, int32_t nx,
                                      const int32_t* xnbr, const void* xnormals, const void* xareas,
                                      const int32_t* x_level_diff, const int32_t* x_offsets, void* stream) {
  try {
    if (!out || (dim != 2 && dim != 3)) return cudaErrorInvalidValue;
    const int    flags = ghost_tail ? 2 : 0;
    t8b200_plan* P     = nullptr;
    const char*  mode_env     = getenv("T8B200_DEVICE_PLAN");
    const bool   generic_only = mode_env && (mode_env[0] == 'g' || mode_env[0] == 's');
    int          rc           = cudaErrorNotSupported;
    if (dim == 3 && !generic_only)   // Subgrid<4,4,4> on structured-only forests: three kernels
      rc = is_f64 ? device_plan_impl<double>(&P, flags, n_local, n_ghost, nf, nb, nbr, (const double*)normals,
                                             (const double*)areas, ranks, indices, nx, xnbr, (const double*)xnormals,
                                             (const double*)xareas, stream, true, (const double*)volumes, level_diff,
                                             x_level_diff)
                  : device_plan_impl<float>(&P, flags, n_local, n_ghost, nf, nb, nbr, (const float*)normals,
                                            (const float*)areas, ranks, indices, nx, xnbr, (const float*)xnormals,
                                            (const float*)xareas, stream, true, (const float*)volumes, level_diff,
                                            x_level_diff);
    if (rc == cudaErrorNotSupported) {   // any other forest: the generic builder over the cell faces
      if (n_local < 0 || n_ghost < 0 || nf < 0 || nb < 0 || nx < 0 || !volumes) return cudaErrorInvalidValue;
      if ((nf + nb > 0 && (!nbr || !normals || !areas)) || (nf > 0 && (!level_diff || !offsets))) return cudaErrorInvalidValue;
      if (nx > 0 && (!xnbr || !xnormals || !xareas || !x_level_diff || !x_offsets)) return cudaErrorInvalidValue;
      if (n_ghost > 0 && (!ranks || !indices)) return cudaErrorInvalidValue;
  #define T8B_SG(TT, DD)                                                                                                    \
    generic_subgrid_plan<TT, DD>(&P, ghost_tail, n_local, n_ghost, nf, nb, nbr, (const TT*)normals, (const TT*)areas,       \
                                 level_diff, offsets, (const TT*)volumes, ranks, indices, nx, xnbr, (const TT*)xnormals,    \
                                 (const TT*)xareas, x_level_diff, x_offsets, stream)
      rc = is_f64 ? (dim == 3 ? T8B_SG(double, 3) : T8B_SG(double, 2)) : (dim == 3 ? T8B_SG(float, 3) : T8B_SG(float, 2));
  #undef T8B_SG
    }
    if (rc) return rc;
    *out = t8b_wrap_subgrid_plan(P, dim);
    return 0;

  } catch (const std::bad_alloc&) {
    return cudaErrorMemoryAllocation;
  } catch (...) {
    cudaGetLastError();
    return cudaErrorUnknown;
  }
}

// test access to every DEVICE array of a plan, numbered as t8b200_plan_host_array does (0 hdr ... 19 blist), as raw bytes in
// the device element type (float plans: area table / normals as float); returns the byte count, -1 on error
int64_t t8b200_plan_device_bytes(const t8b200_plan* P, int which, void* host_out, int64_t capacity_bytes) {
  if (!P || P->host_only) return -1;
  const int64_t nc = P->n_chunks, ts = P->is_f64 ? 8 : 4;
  const void*   src = nullptr;
  int64_t       n   = 0;
  switch (which) {
    case 0: src = P->hdr; n = 32 * nc; break;
    case 1: src = P->halo_elem; n = 4 * nc * P->hs; break;
    case 2: src = P->halo_rank; n = 4 * nc * P->hs; break;
    case 3: src = P->face_lr; n = 4 * nc * P->fs; break;
    case 4: src = P->face_ai; n = nc * P->fs; break;
    case 5: src = P->ell; n = 2 * 8 * std::max<int64_t>(P->n_local, 1); break;
    case 6: src = P->ovf_off; n = 2 * P->n_ovf_off; break;
    case 7: src = P->ovf_ent; n = 2 * P->n_ovf_ent; break;
    case 8: src = P->area_tab; n = ts * P->n_areas; break;
    case 9: src = P->fnx; n = ts * nc * P->fs; break;
    case 10: src = P->fny; n = ts * nc * P->fs; break;
    case 11: src = P->fnz; n = ts * nc * P->fs; break;
    case 12: src = P->farea; n = ts * nc * P->fs; break;
    case 13: src = P->s_rec; n = 16 * (int64_t)P->n_struct; break;
    case 14: src = P->s_halo; n = 1024 * (int64_t)P->n_struct; break;
    case 15: src = P->s_hrank; n = 1024 * (int64_t)P->n_struct; break;
    case 16: src = P->g_list; n = 4 * (int64_t)P->n_generic; break;
    case 17: src = P->pull_rank; n = 4 * P->n_pull; break;
    case 18: src = P->pull_idx; n = 4 * P->n_pull; break;
    case 19: src = P->blist; n = 4 * (int64_t)P->nb_struct; break;
    default: return -1;
  }
  if (!src) return 0;
  if (host_out && n > 0 && cudaMemcpy(host_out, src, (size_t)std::min(n, capacity_bytes), cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  return n;
}

// test access: copies one of the structured / ghost-tail DEVICE arrays of a plan to the host (which: 13 s_rec, 14 s_halo,
// 15 s_hrank, 17 pull_rank, 18 pull_idx, as t8b200_plan_host_array numbers them); returns the element count
int64_t t8b200_plan_device_array(const t8b200_plan* P, int which, int32_t* host_out, int64_t capacity) {
  if (!P || P->host_only) return -1;
  const int32_t* src = nullptr;
  int64_t        n   = 0;
  switch (which) {
    case 13: src = P->s_rec; n = 4 * (int64_t)P->n_struct; break;
    case 14: src = P->s_halo; n = 256 * (int64_t)P->n_struct; break;
    case 15: src = P->s_hrank; n = P->multi ? 256 * (int64_t)P->n_struct : 0; break;
    case 17: src = P->pull_rank; n = P->n_pull; break;
    case 18: src = P->pull_idx; n = P->n_pull; break;
    case 19: src = P->blist; n = P->blist ? P->nb_struct : 0; break;
    default: return -1;
  }
  if (host_out && n > 0 && src) {
    if (cudaMemcpy(host_out, src, sizeof(int32_t) * (size_t)std::min(n, capacity), cudaMemcpyDeviceToHost) != cudaSuccess)
      return -1;
  }
  return src ? n : 0;
}
}
