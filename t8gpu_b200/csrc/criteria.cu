// Refinement indicators of the two example solvers, for sm_100a.  C ABI in include/t8gpu_b200.h (section 6).
//
// Reference behaviour replaced (not translated):
//   examples/compressible_euler/kernels.cu:471-501 estimate_gradient  +  solver.cu:231-245 compute_refinement_criteria:
//     criteria[e] = ( sum over the interior faces f of e of |rho_R - rho_L| ) / cbrt(volume[e]).
//     The reference accumulates with atomicAdd into the Fluxes array (also into other ranks' arrays) and clears it
//     afterwards; here every rank sums the faces of its own elements from the tile plan (which already holds the
//     partition-boundary faces), in a fixed order, without touching the flux array.
//   examples/subgrid/kernels.inl:1109-1168 compute_refinement_criteria<Subgrid>: H1 seminorm of the density over the
//     cells of an element / volume.  One thread walks the 64 (16) cells in the reference's loop order -- the sum is
//     bit-identical to the reference's -- but the densities are staged through shared memory with coalesced loads
//     (the reference's per-thread walk reads 64 strided values per element).
#include "../../include/t8gpu_b200.h"
#include "common.cuh"
#include "tile_plan.cuh"

namespace {

template <typename T>
struct GradArgs {
  const int4*     hdr;
  const int32_t*  halo_elem;
  const int32_t*  halo_rank;
  const uint32_t* face_lr;
  const uint4*    ell;
  const uint16_t* ovf_off;
  const uint16_t* ovf_ent;
  const T*        rho;
  const T* const* rho_all;
  const T*        vol;
  T*              out;
  int64_t         n_local;
  int             hs, fs, split, multi, cmp, vol_shift;
  T               vol_scale;
};

// CTA = one chunk of the tile plan: densities of own + halo elements -> smem, |jump| per face -> smem, per element the
// sum over its faces (table order) / cbrt(volume).
template <typename T>
__global__ void __launch_bounds__(EC) gradient_criteria_kernel(const __grid_constant__ GradArgs<T> A) {
  __shared__ T rho_s[MS];
  __shared__ T jump[MF];
  const int  c = blockIdx.x, tid = threadIdx.x;
  const int4 h0v = __ldg(A.hdr + 2 * c), h1v = __ldg(A.hdr + 2 * c + 1);
  const int  e0c = A.split ? h0v.x : c * EC;
  const int  ecn = A.split ? h0v.y : min(EC, (int)A.n_local - c * EC);
  const int  e = e0c + tid;
  const bool own = tid < ecn;
  const int  nfc = (unsigned)h0v.z >> 16, e2 = A.cmp ? h1v.x : nfc;
  if (own) rho_s[tid] = A.rho[e];
  for (int h = tid; h < A.hs; h += EC) {
    const int idx = A.halo_elem[c * A.hs + h];
    if (idx >= 0) rho_s[EC + h] = A.multi ? A.rho_all[A.halo_rank[c * A.hs + h]][idx] : A.rho[idx];
  }
  __syncthreads();
  for (int j = tid; j < nfc; j += EC) {
    const uint32_t lr = A.face_lr[c * A.fs + j];
    const unsigned sr = lr >> 16;
    // boundary faces do not contribute (the reference loops over the interior faces only, kernels.cu:476-477)
    const bool wall = A.cmp ? j >= e2 : sr == 0xFFFFu;
    jump[j] = wall ? T(0) : t8b200::fabs_(rho_s[sr] - rho_s[lr & 0x3FFFu]);
  }
  __syncthreads();
  if (!own) return;
  const uint4    el = A.ell[e];
  const unsigned w[4] = {el.x, el.y, el.z, el.w};
  T acc = T(0);
#pragma unroll
  for (int s = 0; s < ELL; s++) {
    const unsigned en = (s & 1) ? w[s >> 1] >> 16 : w[s >> 1] & 0xFFFFu;
    if (en != 0xFFFFu) acc += jump[en >> 1];
  }
  if (h1v.y >= 0) {
    const uint16_t* off = A.ovf_off + h1v.y;
    const uint16_t* ent = A.ovf_ent + h1v.z;
    for (int q = off[tid], q1 = off[tid + 1]; q < q1; q++) acc += jump[ent[q] >> 1];
  }
  A.out[e] = acc / cbrt(A.vol[e >> A.vol_shift] * A.vol_scale);   // solver.cu:243
}

// Plans built on the device (device_plan.cu) carry only the halo lists of their structured chunks: the six neighbours
// of an element follow from the box layout, no face records.
template <typename T, class L>
__global__ void __launch_bounds__(256)
structured_gradient_kernel(const int32_t* __restrict__ s_halo, const int32_t* __restrict__ s_hrank, int multi,
                           int my_rank, const T* __restrict__ rho, const T* const* __restrict__ rho_all,
                           const T* __restrict__ vol, T* __restrict__ out) {
  __shared__ T  r[L::NSLOT];
  const int     b = blockIdx.x, t = threadIdx.x;
  const int64_t e = (int64_t)b * 256 + t;
  const int     hidx = s_halo[(int64_t)b * 256 + t];
  const int     hrk  = multi ? s_hrank[(int64_t)b * 256 + t] : my_rank;
  const T       mine = rho[e];
  r[t] = mine;
  r[L::thread_slot(t)] = (multi && hrk != my_rank) ? rho_all[hrk][hidx] : rho[hidx];
  __syncthreads();
  T acc = T(0);
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const int lo = L::at_lower(t, d) ? L::halo_slot(d, 0, L::compact(t, d)) : L::lower_own(t, d);
    const int up = L::at_upper(t, d) ? L::halo_slot(d, 1, L::compact(t, d)) : L::upper_own(t, d);
    acc += t8b200::fabs_(mine - r[lo]);
    acc += t8b200::fabs_(r[up] - mine);
  }
  out[e] = acc / cbrt(vol[e]);   // solver.cu:243
}

// EPB elements per CTA; all threads stage the densities (coalesced), the first EPB threads walk one element each.
template <typename T, int DIM>
__global__ void __launch_bounds__(256)
subgrid_criteria_kernel(int64_t ne, const T* __restrict__ rho, const T* __restrict__ vol, T* __restrict__ out) {
  constexpr int S = DIM == 3 ? 64 : 16, EPB = 32, PAD = S + 1;   // +1: the per-thread walks hit distinct banks
  __shared__ T  d[EPB * PAD];
  const int64_t e0 = (int64_t)blockIdx.x * EPB;
  for (int t = threadIdx.x; t < EPB * S; t += blockDim.x) {
    const int64_t g = e0 * S + t;
    if (g < ne * S) d[(t / S) * PAD + (t % S)] = rho[g];
  }
  __syncthreads();
  const int64_t e = e0 + threadIdx.x;
  if (threadIdx.x >= EPB || e >= ne) return;
  const T* r = d + threadIdx.x * PAD;
  auto     at = [&](int p, int q, int s) { return r[p + 4 * q + 16 * s]; };
  T        h1 = T(0);
  const T  v  = vol[e];
  if (DIM == 3) {
    const T h = cbrt(v) / T(4);   // kernels.inl:1123
    // the reference's three loop nests, in its order; each term is ((a-b)*(a-b))*h added to the running sum
    for (int p = 0; p < 3; p++)
      for (int q = 0; q < 4; q++)
        for (int s = 0; s < 4; s++) { const T x = at(p + 1, q, s) - at(p, q, s); h1 += x * x * h; }
    for (int p = 0; p < 4; p++)
      for (int q = 0; q < 3; q++)
        for (int s = 0; s < 4; s++) { const T x = at(p, q + 1, s) - at(p, q, s); h1 += x * x * h; }
    for (int p = 0; p < 4; p++)
      for (int q = 0; q < 4; q++)
        for (int s = 0; s < 3; s++) { const T x = at(p, q, s + 1) - at(p, q, s); h1 += x * x * h; }
  } else {
    const T h = sqrt(v) / T(4);   // kernels.inl:1152
    for (int p = 0; p < 3; p++)
      for (int q = 0; q < 4; q++) { const T x = at(p + 1, q, 0) - at(p, q, 0); h1 += x * x * h; }
    for (int p = 0; p < 4; p++)
      for (int q = 0; q < 3; q++) { const T x = at(p, q + 1, 0) - at(p, q, 0); h1 += x * x * h; }
  }
  out[e] = h1 / v;
}

template <typename T>
int gradient_impl(const t8b200_plan* P, const T* rho, const T* const* rho_all, const T* vol, T* out, void* stream) {
  if (!P || P->host_only || !rho || !vol || !out || (P->multi && !P->ghost_tail && !rho_all)) return cudaErrorInvalidValue;
  if (P->is_f64 != (sizeof(T) == 8)) return cudaErrorInvalidValue;
  if (P->n_chunks == 0) return 0;
  if (!P->hdr) {   // device-built plan: every chunk is a structured box of hexahedra
    if (P->n_struct != P->n_chunks || P->box_layout != 0 || P->vol_shift != 0) return cudaErrorInvalidValue;
    structured_gradient_kernel<T, t8b200::MortonBox><<<P->n_chunks, 256, 0, (cudaStream_t)stream>>>(
        P->s_halo, P->s_hrank, P->multi && !P->ghost_tail, P->my_rank, rho, rho_all, vol, out);
    return cudaGetLastError();
  }
  GradArgs<T> A{};
  A.hdr = reinterpret_cast<const int4*>(P->hdr);
  A.halo_elem = P->halo_elem; A.halo_rank = P->halo_rank; A.face_lr = P->face_lr;
  A.ell = P->ell; A.ovf_off = P->ovf_off; A.ovf_ent = P->ovf_ent;
  A.rho = rho; A.rho_all = rho_all; A.vol = vol; A.out = out; A.n_local = P->n_local;
  A.hs = P->hs; A.fs = P->fs; A.split = P->split; A.multi = P->multi && !P->ghost_tail; A.cmp = P->cmp;
  A.vol_shift = P->vol_shift; A.vol_scale = (T)P->vol_scale;
  gradient_criteria_kernel<T><<<P->n_chunks, EC, 0, (cudaStream_t)stream>>>(A);
  return cudaGetLastError();
}

template <typename T>
int subgrid_impl(int dim, int64_t ne, const T* rho, const T* vol, T* out, void* stream) {
  if ((dim != 2 && dim != 3) || ne < 0) return cudaErrorInvalidValue;
  if (ne == 0) return 0;
  if (!rho || !vol || !out) return cudaErrorInvalidValue;
  const unsigned blocks = (unsigned)((ne + 31) / 32);
  if (dim == 3) subgrid_criteria_kernel<T, 3><<<blocks, 256, 0, (cudaStream_t)stream>>>(ne, rho, vol, out);
  else subgrid_criteria_kernel<T, 2><<<blocks, 256, 0, (cudaStream_t)stream>>>(ne, rho, vol, out);
  return cudaGetLastError();
}
}  // namespace

extern "C" {
int t8b200_gradient_criteria_f32(const t8b200_plan* plan, const float* rho, const float* const* rho_all,
                                 const float* vol, float* criteria, void* stream) {
  return gradient_impl<float>(plan, rho, rho_all, vol, criteria, stream);
}
int t8b200_gradient_criteria_f64(const t8b200_plan* plan, const double* rho, const double* const* rho_all,
                                 const double* vol, double* criteria, void* stream) {
  return gradient_impl<double>(plan, rho, rho_all, vol, criteria, stream);
}
int t8b200_subgrid_criteria_f32(int dim, int64_t n_elements, const float* rho, const float* vol, float* criteria,
                                void* stream) {
  return subgrid_impl<float>(dim, n_elements, rho, vol, criteria, stream);
}
int t8b200_subgrid_criteria_f64(int dim, int64_t n_elements, const double* rho, const double* vol, double* criteria,
                                void* stream) {
  return subgrid_impl<double>(dim, n_elements, rho, vol, criteria, stream);
}
}
