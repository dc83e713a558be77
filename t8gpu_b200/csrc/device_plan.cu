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
#include <thrust/sort.h>
#include <thrust/unique.h>

#include <cmath>

#include "../../include/t8gpu_b200.h"
#include "box_layout.cuh"
#include "common.cuh"
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
  if (i >= n || hrank[i] == me) return;
  keys[atomicAdd(count, 1ull)] = ((unsigned long long)(uint32_t)hrank[i] << 32) | (uint32_t)halo[i];
}
__global__ void redirect_kernel(int64_t n, int32_t* __restrict__ halo, int32_t* __restrict__ hrank, int me,
                                const unsigned long long* __restrict__ keys, int64_t nkeys, int64_t n_local) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || hrank[i] == me) return;
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

}  // namespace

extern "C" {
int t8b200_plan_create_device(t8b200_plan** out, int is_f64, int ghost_tail, int64_t n_local, int64_t n_ghost,
                              int32_t nf, int32_t nb, const int32_t* nbr, const void* normals, const void* areas,
                              const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                              const void* xnormals, const void* xareas, void* stream) {
  const int flags = ghost_tail ? 2 : 0;
  if (is_f64)
    return device_plan_impl<double>(out, flags, n_local, n_ghost, nf, nb, nbr, (const double*)normals,
                                    (const double*)areas, ranks, indices, nx, xnbr, (const double*)xnormals,
                                    (const double*)xareas, stream);
  return device_plan_impl<float>(out, flags, n_local, n_ghost, nf, nb, nbr, (const float*)normals, (const float*)areas,
                                 ranks, indices, nx, xnbr, (const float*)xnormals, (const float*)xareas, stream);
}

int t8b200_subgrid_plan_create_device(t8b200_subgrid_plan** out, int is_f64, int dim, int ghost_tail, int64_t n_local,
                                      int64_t n_ghost, int32_t nf, int32_t nb, const int32_t* nbr, const void* normals,
                                      const void* areas, const int32_t* level_diff, const void* volumes,
                                      const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                                      const void* xnormals, const void* xareas, const int32_t* x_level_diff,
                                      void* stream) {
  if (!out) return cudaErrorInvalidValue;
  if (dim != 3) return cudaErrorNotSupported;     // Subgrid<4,4>: host builder
  const int    flags = ghost_tail ? 2 : 0;
  t8b200_plan* P     = nullptr;
  const int rc = is_f64 ? device_plan_impl<double>(&P, flags, n_local, n_ghost, nf, nb, nbr, (const double*)normals,
                                                   (const double*)areas, ranks, indices, nx, xnbr, (const double*)xnormals,
                                                   (const double*)xareas, stream, true, (const double*)volumes, level_diff,
                                                   x_level_diff)
                        : device_plan_impl<float>(&P, flags, n_local, n_ghost, nf, nb, nbr, (const float*)normals,
                                                  (const float*)areas, ranks, indices, nx, xnbr, (const float*)xnormals,
                                                  (const float*)xareas, stream, true, (const float*)volumes, level_diff,
                                                  x_level_diff);
  if (rc) return rc;
  *out = t8b_wrap_subgrid_plan(P, 3);
  return 0;
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
