// Fused tile-plan RK stage for unstructured (MeshManager) meshes on sm_100a: face flux + per-element accumulation +
// SSP-RK3 combination + wave-speed reduction in ONE kernel per stage; the flux accumulators never reach HBM.
// C ABI in include/t8gpu_b200.h (section 2).
//
// Reference behaviour replaced (not translated): one stage of CompressibleEulerSolver::iterate,
//   examples/compressible_euler/solver.cu:78-112 = kepes_compute_fluxes (kernels.cu:135-309) +
//   reflective_boundary_condition (kernels.cu:311-469) + SSP_3RK_step{1,2,3} (ssp_runge_kutta.inl:30-99),
//   and the thrust::reduce of solver.cu:213-217.
//
// The plan (connectivity re-laid out per chunk) is described and built in tile_plan.cuh.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../include/t8gpu_b200.h"
#include "common.cuh"
#include "euler_flux.cuh"
#include "mesh_faces.cuh"
#include "peer_sync.cuh"
#include "plan_emulate.cuh"
#include "tile_plan.cuh"

using namespace t8b200;

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
  int             wave;  // CTAs resident on the device at once = distance of the next-wave prefetch (0: off)
  const char*     pf_ptr[16];   // L2 prefetch table: base pointer (null: skip) and bytes per index
  unsigned        pf_unit[16];
  const int32_t*  chunk_list;   // SPLIT variants: chunk id per CTA (plans whose structured chunks run in structured.cu)
  const uint4*    ell;
  const uint16_t* ovf_off;
  const uint16_t* ovf_ent;
  const T*        in[5];
  const T* const* in_all[5];
  const T*        prev[5];
  T*              out[5];
  const T*        vol;
  int             vol_shift;
  T               vol_scale;
  T               dt;
  const T*        dt_ptr;       // non-null: the time step is read from device memory (t8b200_timestep_*)
  StageSync       sync;         // multi-GPU stage ordering done by the kernel itself (mailboxes == nullptr: off)
  T*              speed_max;
  int64_t         n_local;
  int             stage;
  int             multi;
  int             my_rank;
};

// Shared-memory layout of the staged data.
//   fp64: struct of arrays, cells [7][MS], fluxes [5][MF] (64-bit accesses, conflict-free for neighbouring slots).
//   fp32: the kernel is issue-bound, so the cells are two float4 records per slot (rho,hx,hy,hz | kp,b,q,-) and the
//         fluxes one float4 + one float per face: 2 instead of 7 loads per cell, 2 instead of 5 per flux.
template <typename T, int MS, int MF>
struct Smem;
template <int MS, int MF>
struct Smem<double, MS, MF> {
  static constexpr size_t bytes = sizeof(double) * ((size_t)NCELLQ * MS + 5 * (size_t)MF);
  double* cq;
  double* fl;
  __device__ explicit Smem(unsigned char* raw) : cq(reinterpret_cast<double*>(raw)), fl(cq + NCELLQ * MS) {}
  __device__ __forceinline__ void store_cell(int s, const Cell<double>& q) const {
    cq[0 * MS + s] = q.rho; cq[1 * MS + s] = q.hx; cq[2 * MS + s] = q.hy; cq[3 * MS + s] = q.hz;
    cq[4 * MS + s] = q.kp;  cq[5 * MS + s] = q.b;  cq[6 * MS + s] = q.q;
  }
  __device__ __forceinline__ Cell<double> load_cell(int s) const {
    Cell<double> q;
    q.rho = cq[0 * MS + s]; q.hx = cq[1 * MS + s]; q.hy = cq[2 * MS + s]; q.hz = cq[3 * MS + s];
    q.kp  = cq[4 * MS + s]; q.b  = cq[5 * MS + s]; q.q  = cq[6 * MS + s];
    return q;
  }
  __device__ __forceinline__ void store_flux(int j, const double F[5]) const {
#pragma unroll
    for (int k = 0; k < 5; k++) fl[k * MF + j] = F[k];
  }
  // Axis-permuted access for faces with normal +e_a: (p0,p1,p2) = (a, a+1, a+2) mod 3 is a cyclic permutation of the
  // axes, i.e. a rotation, under which the flux is invariant; reading the velocity rows / writing the momentum-flux
  // rows in that order lets ONE copy of the x-normal flux serve all three axes (the rows are only addresses here).
  __device__ __forceinline__ Cell<double> load_cell_axis(int s, int p0, int p1, int p2) const {
    Cell<double> q;
    q.rho = cq[0 * MS + s]; q.hx = cq[(1 + p0) * MS + s]; q.hy = cq[(1 + p1) * MS + s]; q.hz = cq[(1 + p2) * MS + s];
    q.kp  = cq[4 * MS + s]; q.b  = cq[5 * MS + s]; q.q  = cq[6 * MS + s];
    return q;
  }
  __device__ __forceinline__ void store_flux_axis(int j, const double F[5], int p0, int p1, int p2) const {
    fl[j] = F[0]; fl[(1 + p0) * MF + j] = F[1]; fl[(1 + p1) * MF + j] = F[2]; fl[(1 + p2) * MF + j] = F[3];
    fl[4 * MF + j] = F[4];
  }
  __device__ __forceinline__ void zero_flux(int j) const {
#pragma unroll
    for (int k = 0; k < 5; k++) fl[k * MF + j] = 0.0;
  }
  // one entry of the element -> face table: acc += (side ? +1 : -1) * flux[face]; an empty entry (0xFFFF) reads the
  // zero column MF - 1 (the plan never uses it), so the entries need no branches and their loads overlap
  __device__ __forceinline__ void gather(unsigned en, double acc[5]) const {
    const int    j  = min((int)(en >> 1), MF - 1);
    const double sg = (en & 1u) ? 1.0 : -1.0;
#pragma unroll
    for (int k = 0; k < 5; k++) acc[k] = fma(sg, fl[k * MF + j], acc[k]);
  }
};
template <int MS, int MF>
struct Smem<float, MS, MF> {
  static constexpr size_t bytes = sizeof(float4) * 2 * (size_t)MS + sizeof(float4) * (size_t)MF + sizeof(float) * (size_t)MF;
  float4* ca;   // [MS] rho, hx, hy, hz
  float4* cb;   // [MS] kp, b, q, -
  float4* f4;   // [MF] F0..F3
  float*  f1;   // [MF] F4
  __device__ explicit Smem(unsigned char* raw)
      : ca(reinterpret_cast<float4*>(raw)), cb(ca + MS), f4(cb + MS), f1(reinterpret_cast<float*>(f4 + MF)) {}
  __device__ __forceinline__ void store_cell(int s, const Cell<float>& q) const {
    ca[s] = make_float4(q.rho, q.hx, q.hy, q.hz);
    cb[s] = make_float4(q.kp, q.b, q.q, 0.f);
  }
  __device__ __forceinline__ Cell<float> load_cell(int s) const {
    const float4 a = ca[s], b = cb[s];
    Cell<float>  q;
    q.rho = a.x; q.hx = a.y; q.hy = a.z; q.hz = a.w; q.kp = b.x; q.b = b.y; q.q = b.z;
    return q;
  }
  __device__ __forceinline__ void store_flux(int j, const float F[5]) const {
    f4[j] = make_float4(F[0], F[1], F[2], F[3]);
    f1[j] = F[4];
  }
  // records are vectors here, so the permutation is done in registers
  __device__ __forceinline__ Cell<float> load_cell_axis(int s, int p0, int, int) const {
    const Cell<float> c = load_cell(s);
    Cell<float>       q = c;
    q.hx = p0 == 0 ? c.hx : p0 == 1 ? c.hy : c.hz;
    q.hy = p0 == 0 ? c.hy : p0 == 1 ? c.hz : c.hx;
    q.hz = p0 == 0 ? c.hz : p0 == 1 ? c.hx : c.hy;
    return q;
  }
  __device__ __forceinline__ void store_flux_axis(int j, const float F[5], int p0, int, int) const {
    f4[j] = make_float4(F[0], p0 == 0 ? F[1] : p0 == 1 ? F[3] : F[2], p0 == 0 ? F[2] : p0 == 1 ? F[1] : F[3],
                        p0 == 0 ? F[3] : p0 == 1 ? F[2] : F[1]);
    f1[j] = F[4];
  }
  __device__ __forceinline__ void zero_flux(int j) const {
    f4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    f1[j] = 0.f;
  }
  __device__ __forceinline__ void gather(unsigned en, float acc[5]) const {
    const int    j  = min((int)(en >> 1), MF - 1);
    const float  sg = (en & 1u) ? 1.f : -1.f;
    const float4 a  = f4[j];
    acc[0] = fmaf(sg, a.x, acc[0]); acc[1] = fmaf(sg, a.y, acc[1]); acc[2] = fmaf(sg, a.z, acc[2]);
    acc[3] = fmaf(sg, a.w, acc[3]); acc[4] = fmaf(sg, f1[j], acc[4]);
  }
};

// bulk L2 prefetch of [p, p + bytes): 16-byte granules, address aligned down (a hint; used on the state rows inside
// [0, n_local) and on plan arrays, which upload() pads)
__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
  const unsigned long long a = (unsigned long long)p & ~15ull;
  bytes = (bytes + (unsigned)((unsigned long long)p & 15ull)) & ~15u;
  if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes) : "memory");
}

#ifdef T8B_PHASE_CLOCKS   // tools/build_variant.py: per-phase CTA timing (sum of clock64 deltas of warp 0 and warp 7)
__device__ unsigned long long t8b_phase_clk[16];
__device__ long long t8b_cta_log[4 * 65536];   // per CTA (index < 65536): smid, clock at start, at phase 1, at end
extern "C" int t8b200_debug_cta_log(long long* out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, t8b_cta_log, sizeof(t8b_cta_log));
}
#define T8B_CLK(i) do { if ((tid & 31) == 0 && (tid == 0 || tid == 224)) clk[i] = clock64(); } while (0)
extern "C" int t8b200_debug_phase_clocks(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, t8b_phase_clk, sizeof(t8b_phase_clk));
  if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(t8b_phase_clk, z, sizeof(z)); }
  return 0;
}
#else
#define T8B_CLK(i)
#endif

// same for an address known to be 16-byte aligned (the size is rounded down to the granule)
__device__ __forceinline__ void prefetch_l2_aligned(const void* p, unsigned bytes) {
  bytes &= ~15u;
  if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// CTA = one chunk of EC consecutive elements.  MS / MF: compile-time strides of the shared-memory SoA arrays
// (slots = EC own + halo; faces), so every shared access is base + index*sizeof(T) + immediate.
//   phase 0: conserved -> per-cell quantities for the chunk's own elements (coalesced) and its halo (gather)
//   phase 1: every face touching the chunk: flux from the staged cells -> smem
//   phase 2: per element: signed gather of its faces' fluxes (fixed order: deterministic), RK combination, store
template <typename T, int MS, int MF, int MINB, bool CMP, bool SPLIT, bool SMAX>
__global__ void __launch_bounds__(EC, MINB) fused_stage_kernel(const __grid_constant__ FusedArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Smem<T, MS, MF> sm(smem_raw);
  __shared__ T atab[CMP ? 256 : 1];
  __shared__ T red[EC / 32];
  // chunk list entries: chunk id, bit 30 = partition-boundary chunk (tile_plan.cuh)
  const int cl  = (SPLIT && A.chunk_list) ? __ldg(A.chunk_list + blockIdx.x) : (int)blockIdx.x;
  const int c   = cl & 0x3FFFFFFF;
  const int tid = threadIdx.x;
#ifdef T8B_PHASE_CLOCKS
  long long clk[7] = {0, 0, 0, 0, 0, 0, 0};
#endif
  T8B_CLK(0);

  // ---- phase 0: issue every independent global load of the prologue first.  The halo indices and the first face
  // record sit at fixed strides, so nothing here waits for the chunk header.
  const int4 h0v = __ldg(A.hdr + 2 * c), h1v = __ldg(A.hdr + 2 * c + 1);
  // chunk c covers elements [c * EC, ...) unless blocks were split (SPLIT: the header says, and only then do the
  // loads of the own elements wait for it).  Element, halo and face indices fit 32 bits (checked by the plan).
  const int e0c = SPLIT ? h0v.x : c * EC;
  const int ecn = SPLIT ? h0v.y : min(EC, (int)A.n_local - c * EC);
  const int  e   = e0c + tid;
  const bool own = tid < ecn;
  const int  hb = c * A.hs, fb = c * A.fs;
  T   u0 = T(1), u1 = T(0), u2 = T(0), u3 = T(0), u4 = T(1);
  int hidx = -1, hrk = 0;
  if (tid < A.hs) {
    hidx = A.halo_elem[hb + tid];
    if (A.multi) hrk = A.halo_rank[hb + tid];
  }
#ifdef T8B_ABLATE_MEM   // timing experiment only (compute floor: no state traffic), results are wrong
  u0 = T(1) + T(1e-3) * T(tid); u1 = T(0.1); u2 = T(0.2); u3 = T(0.3); u4 = T(3) + T(1e-3) * T(tid & 7);
#else
  if (own) {
    u0 = A.in[0][e]; u1 = A.in[1][e]; u2 = A.in[2][e]; u3 = A.in[3][e]; u4 = A.in[4][e];
  }
#endif
  // first face record of this thread (consumed in phase 1; its latency hides behind phase 0)
  uint32_t lr_n = 0;
  if (tid < A.fs) lr_n = A.face_lr[fb + tid];
  // fp32 (issue-bound): the warp index is taken through a shuffle from lane 0, which the compiler knows to be
  // warp-uniform, so the hints are computed on the uniform datapath by the whole warp without the per-lane loop it
  // otherwise wraps around a uniform instruction (-2.2 % per step); fp64 measured 1.2 % faster with lane 0 alone.
  constexpr bool UNIPF = sizeof(T) == 4;
  if ((UNIPF || (tid & 31) == 0) && tid < 256) {
    // L2 prefetch hints, two per warp (a bulk prefetch is a warp-uniform instruction: spreading them avoids a serial
    // loop in one warp), table-driven (A.pf_*, filled by the host).  Items 0-7: the phase-2 operands of this chunk, so
    // that their loads before the barrier are L2 hits.  Items 8-15 (SPLIT == false only: addresses computable
    // without a header): the streams of the chunk that takes over a CTA slot about one wave later (CTAs are
    // dispatched in index order): 8-12 state rows, 13-15 halo indices, face records, header.
    const int w  = UNIPF ? __shfl_sync(0xffffffffu, tid >> 5, 0) : tid >> 5;   // 0..7
    const int cw = c + A.wave;
    // (unsplit plans: chunk offsets are multiples of EC elements and the host only fills the table with 16-byte aligned
    //  rows, so the address needs no alignment arithmetic)
    if (A.pf_ptr[w]) {
      if (SPLIT) prefetch_l2(A.pf_ptr[w] + (size_t)e0c * A.pf_unit[w], (unsigned)ecn * A.pf_unit[w]);
      else prefetch_l2_aligned(A.pf_ptr[w] + (size_t)e0c * A.pf_unit[w], (unsigned)ecn * A.pf_unit[w]);
    }
    if (!SPLIT && A.pf_ptr[w + 8] && cw < (int)gridDim.x) {
      const unsigned idx = w < 5 ? (unsigned)cw * EC : (unsigned)cw;
      const unsigned cnt = w < 5 ? (unsigned)min(EC, (int)A.n_local - cw * EC) : 1u;
      prefetch_l2_aligned(A.pf_ptr[w + 8] + (size_t)idx * A.pf_unit[w + 8], cnt * A.pf_unit[w + 8]);
    }
  }
  if (CMP && tid < A.n_areas) atab[tid] = A.area_tab[tid];
  if (tid == EC - 1) sm.zero_flux(MF - 1);

  auto load_halo = [&](int idx, int rk, T& a0, T& a1, T& a2, T& a3, T& a4) {
    if (A.multi && rk != A.my_rank) {   // ghost: through the [var][rank] tables (a peer GPU's array over NVLink)
      // self-ordering launches: the owner's previous stage must be complete before its element is read
      if (A.sync.mailboxes != nullptr) stage_wait_owner(A.sync, rk);
      a0 = A.in_all[0][rk][idx]; a1 = A.in_all[1][rk][idx]; a2 = A.in_all[2][rk][idx];
      a3 = A.in_all[3][rk][idx]; a4 = A.in_all[4][rk][idx];
    } else {
#ifdef T8B_ABLATE_MEM
      a0 = T(1) + T(1e-3) * T(idx & 255); a1 = T(0.1); a2 = T(0.2); a3 = T(0.3); a4 = T(3) + T(1e-3) * T(idx & 7);
#else
      a0 = A.in[0][idx]; a1 = A.in[1][idx]; a2 = A.in[2][idx]; a3 = A.in[3][idx]; a4 = A.in[4][idx];
#endif
    }
  };
  auto convert_halo = [&](int h, int idx, int rk) {
    T a0, a1, a2, a3, a4;
    load_halo(idx, rk, a0, a1, a2, a3, a4);
    sm.store_cell(EC + h, to_cell(a0, a1, a2, a3, a4));
  };
  // partition-boundary chunk of a launch that orders itself against the peers: it signals when it is done
  const bool bnd = SPLIT && A.sync.mailboxes != nullptr && (cl >> 30) != 0;
  // the gathers of the halo states are issued before the own elements are converted: the two latencies overlap
  T g0 = T(1), g1 = T(0), g2 = T(0), g3 = T(0), g4 = T(1);
  if (hidx >= 0) load_halo(hidx, hrk, g0, g1, g2, g3, g4);
  if (own) sm.store_cell(tid, to_cell(u0, u1, u2, u3, u4));
  if (hidx >= 0) sm.store_cell(EC + tid, to_cell(g0, g1, g2, g3, g4));
  if (A.hs > EC) {   // only plans with more than EC halo entries in some chunk (adaptive meshes)
    for (int h = tid + EC; h < A.hs; h += EC) {
      const int idx = A.halo_elem[hb + h];
      if (idx >= 0) convert_halo(h, idx, A.multi ? A.halo_rank[hb + h] : 0);
    }
  }
  const int  nfc = (unsigned)h0v.z >> 16;
  const int  e0 = h0v.w & 0xFFFF, e1 = (unsigned)h0v.w >> 16, e2 = h1v.x;
  const int  area_idx  = h1v.w;
  constexpr bool want_smax = SMAX;   // stage 3 with a CFL reduction; a template parameter frees its registers elsewhere
  T8B_CLK(1);
  __syncthreads();
  T8B_CLK(2);

  // ---- phase 1: software-pipelined over this thread's faces (record j+EC is in flight while j is evaluated)
  T smax = T(0);
#define T8B_SMAX_UPDATE(s) do { if (want_smax) smax = fmax_(smax, (s)); } while (0)
#ifndef T8B_AXPERM
#define T8B_AXPERM (sizeof(T) == 8)
#endif
  constexpr bool AXPERM = T8B_AXPERM;
  if (CMP) {
    // SCALE = false: every face of the chunk has the same area, applied once per element in phase 2
    auto interior = [&](auto scale_tag) {
      constexpr bool SCALE = decltype(scale_tag)::value;
      for (int j = tid; j < e2; j += EC) {
        const uint32_t lr = lr_n;
        if (j + EC < nfc) lr_n = A.face_lr[fb + j + EC];
        T F[5], s;
        if (AXPERM) {   // one copy of the x-normal flux, the axis enters through the row addresses
          const int p0 = (lr >> 14) & 3, p1 = p0 == 2 ? 0 : p0 + 1, p2 = p0 == 0 ? 2 : p0 - 1;   // axis from the record
          const Cell<T> L = sm.load_cell_axis(lr & 0x3FFFu, p0, p1, p2);
          const Cell<T> R = sm.load_cell_axis(lr >> 16, p0, p1, p2);
          s = kepes_flux_n<T, 0>(L, R, T(0), T(0), T(0), F);
          T8B_SMAX_UPDATE(s);
          if (SCALE) {
            const T ar = atab[A.face_ai[fb + j]];
#pragma unroll
            for (int k = 0; k < 5; k++) F[k] *= ar;
          }
          sm.store_flux_axis(j, F, p0, p1, p2);
          continue;
        }
        const Cell<T> L = sm.load_cell(lr & 0x3FFFu);
        const Cell<T> R = sm.load_cell(lr >> 16);
#ifdef T8B_ABLATE_FLUX   // tools/build_variant.py: timing experiment only (memory + indexing floor), results are wrong
        F[0] = L.rho - R.rho; F[1] = L.hx - R.hx; F[2] = L.hy - R.hy; F[3] = L.hz - R.hz; F[4] = L.kp - R.kp + L.b - R.b + L.q - R.q;
        s = F[0];
#else
        if (j < e0) s = kepes_flux_n<T, 0>(L, R, T(0), T(0), T(0), F);
        else if (j < e1) s = kepes_flux_n<T, 1>(L, R, T(0), T(0), T(0), F);
        else s = kepes_flux_n<T, 2>(L, R, T(0), T(0), T(0), F);
#endif
        T8B_SMAX_UPDATE(s);
        if (SCALE) {
          const T ar = atab[A.face_ai[fb + j]];
#pragma unroll
          for (int k = 0; k < 5; k++) F[k] *= ar;
        }
        sm.store_flux(j, F);
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
      const Cell<T> L = sm.load_cell(lr & 0x3FFFu);
      const Cell<T> R = mirror(L, nx, ny, nz);
      T F[5];
      const T s = kepes_flux_n<T, -1>(L, R, nx, ny, nz, F);
      T8B_SMAX_UPDATE(s);
      if (area_idx < 0) {
        const T ar = atab[A.face_ai[fb + j]];
#pragma unroll
        for (int k = 0; k < 5; k++) F[k] *= ar;
      }
      sm.store_flux(j, F);
    }
  } else {
    T nx_n = T(0), ny_n = T(0), nz_n = T(0), ar_n = T(0);
    if (tid < nfc) { nx_n = A.fnx[fb + tid]; ny_n = A.fny[fb + tid]; nz_n = A.fnz[fb + tid]; ar_n = A.farea[fb + tid]; }
    for (int j = tid; j < nfc; j += EC) {
      const uint32_t lr = lr_n;
      const T        nx = nx_n, ny = ny_n, nz = nz_n, ar = ar_n;
      if (j + EC < nfc) {
        const int g = fb + j + EC;
        lr_n = A.face_lr[g]; nx_n = A.fnx[g]; ny_n = A.fny[g]; nz_n = A.fnz[g]; ar_n = A.farea[g];
      }
      const int     sr = lr >> 16;
      const Cell<T> L  = sm.load_cell(lr & 0x3FFFu);
      const Cell<T> R  = sr == 0xFFFF ? mirror(L, nx, ny, nz) : sm.load_cell(sr);
      T F[5];
      const T s = kepes_flux_n<T, -1>(L, R, nx, ny, nz, F);
      T8B_SMAX_UPDATE(s);
#pragma unroll
      for (int k = 0; k < 5; k++) F[k] *= ar;
      sm.store_flux(j, F);
    }
  }

  T8B_CLK(6);
  // ---- phase 2: the element->face table, the volume and the old states are requested BEFORE the barrier so their
  // latency overlaps the wait.  The element's own conserved values are read again (L1/L2 hit) instead of being
  // carried in 10 registers through the face loop, and combined with U^n right away (5 live values, not 10).
  const int stage = A.stage;
  uint4     el = make_uint4(~0u, ~0u, ~0u, ~0u);
  T         vol = T(1), base[5] = {T(0), T(0), T(0), T(0), T(0)};
  if (own) {
    el  = A.ell[e];
#ifdef T8B_ABLATE_MEM
    vol = A.vol_scale;
#pragma unroll
    for (int k = 0; k < 5; k++) base[k] = T(k + tid);
#else
    vol = A.vol[e >> A.vol_shift] * A.vol_scale;
#pragma unroll
    for (int k = 0; k < 5; k++) base[k] = A.in[k][e];
    if (stage != 1) {
      const T cp = stage == 2 ? T(0.75) : T(0.33333333333333), ci = stage == 2 ? T(0.25) : T(0.66666666666666);
#pragma unroll
      for (int k = 0; k < 5; k++) base[k] = cp * A.prev[k][e] + ci * base[k];
    }
#endif
  }
  // the overflow-table words of the header are read again here rather than kept in registers through the face loops
  // (fp32, 48 registers: frees two and removes its spills, -2.6 % per step; fp64 measured 0.5 % faster without)
  int2 ovf = make_int2(h1v.y, h1v.z);
  if (sizeof(T) == 4) {
    const int4 h1b = __ldg(A.hdr + 2 * c + 1);
    ovf = make_int2(h1b.y, h1b.z);
  }
  T sc = fast_rcp(vol) * (A.dt_ptr ? __ldg(A.dt_ptr) : A.dt);   // before the barrier: the reciprocal chain hides in the wait
  if (stage == 2) sc *= T(0.25);
  if (stage == 3) sc *= T(0.66666666666666);
  T8B_CLK(3);
  __syncthreads();
  T8B_CLK(4);

  if (own) {
    T acc[5] = {T(0), T(0), T(0), T(0), T(0)};
    const unsigned w[4] = {el.x, el.y, el.z, el.w};
#pragma unroll
    for (int s = 0; s < ELL; s++) {
      const unsigned en = (s & 1) ? w[s >> 1] >> 16 : w[s >> 1] & 0xFFFFu;
      // fp32: the first six entries without a branch (empty ones read the zero column), so their loads overlap:
      // -4 % per step; fp64: the 60 registers the overlapped loads would need are not there, branches measured faster
      if (s < (sizeof(T) == 4 ? 6 : 0) || en != 0xFFFFu) sm.gather(en, acc);
    }
    if (ovf.x >= 0) {  // rare: elements of this chunk with more than ELL faces
      const uint16_t* off = A.ovf_off + ovf.x;
      const uint16_t* ent = A.ovf_ent + ovf.y;
#pragma unroll 1
      for (int q = off[tid], q1 = off[tid + 1]; q < q1; q++) sm.gather(ent[q], acc);
    }
    if (CMP && area_idx >= 0) sc *= atab[area_idx];
#ifdef T8B_ABLATE_MEM
    if (acc[0] + acc[1] + acc[2] + acc[3] + acc[4] == T(123.456))
#endif
#pragma unroll
    for (int k = 0; k < 5; k++) A.out[k][e] = base[k] + sc * acc[k];
  }

#ifdef T8B_PHASE_CLOCKS
  T8B_CLK(5);
  if (tid == 0 || tid == 224) {
    const int o = tid == 0 ? 0 : 8;
    for (int i = 0; i < 5; i++) atomicAdd(&t8b_phase_clk[o + i], (unsigned long long)(clk[i + 1] - clk[i]));
    atomicAdd(&t8b_phase_clk[o + 5], 1ull);
    atomicAdd(&t8b_phase_clk[o + 6], (unsigned long long)(clk[6] - clk[2]));   // the face loops alone
    if (tid == 0 && c < 65536) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      t8b_cta_log[4 * c] = smid; t8b_cta_log[4 * c + 1] = clk[0]; t8b_cta_log[4 * c + 2] = clk[2]; t8b_cta_log[4 * c + 3] = clk[5];
    }
  }
#endif
  if (SMAX) {
    smax = warp_max(smax);
    if ((tid & 31) == 0) red[tid >> 5] = smax;
    __syncthreads();
    if (tid == 0) {
      T m = red[0];
      for (int w = 1; w < EC / 32; w++) m = fmax_(m, red[w]);
      atomic_max_nonneg(A.speed_max, m);
    }
  }
  if (bnd && A.sync.signal_epoch > 0) {   // every store of this chunk precedes the count (and the flag behind it)
    __syncthreads();
    if (tid == 0) stage_signal(A.sync);
  }
}

template <typename T, int MS, int MF, int MINB, bool CMP, bool SPLIT, bool SMAX>
static int launch_variant(const t8b200_plan* P, const FusedArgs<T>& A, cudaStream_t st) {
  auto k = fused_stage_kernel<T, MS, MF, MINB, CMP, SPLIT, SMAX>;
#ifdef T8B_PHASE_CLOCKS   // analysis builds: T8B200_SMEM_PAD lowers the occupancy (e.g. 140000: one CTA per SM)
  static const size_t smem = Smem<T, MS, MF>::bytes + (getenv("T8B200_SMEM_PAD") ? atoi(getenv("T8B200_SMEM_PAD")) : 0);
#else
  constexpr size_t smem = Smem<T, MS, MF>::bytes;
#endif
  // CTAs of this variant the device holds at once; the opt-in above 48 KB of shared memory is per device, so it is
  // set and cached per device id (a process-wide flag broke a second device / a device reset)
  static int resident[64];
  int        dev = 0;
  T8B_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (resident[dev] == 0) {
    T8B_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0, per_sm = 0;
    T8B_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    T8B_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, EC, smem));
    resident[dev] = std::max(1, sms * per_sm);
  }
  FusedArgs<T> B = A;
  static const int wave_knob = getenv("T8B200_WAVE") ? atoi(getenv("T8B200_WAVE")) : -1;   // tuning knob
  B.wave = wave_knob >= 0 ? wave_knob : resident[dev];
  for (int k = 0; k < 16; k++) { B.pf_ptr[k] = nullptr; B.pf_unit[k] = 0; }
  for (int k = 0; k < 5; k++) {
    if (B.stage != 1) { B.pf_ptr[k] = (const char*)B.prev[k]; B.pf_unit[k] = sizeof(T); }
    if (B.wave > 0) { B.pf_ptr[8 + k] = (const char*)B.in[k]; B.pf_unit[8 + k] = sizeof(T); }
  }
  if (B.vol_shift == 0) { B.pf_ptr[5] = (const char*)B.vol; B.pf_unit[5] = sizeof(T); }
  B.pf_ptr[6] = (const char*)B.ell; B.pf_unit[6] = sizeof(uint4);
  if (B.wave > 0) {
    B.pf_ptr[13] = (const char*)B.halo_elem; B.pf_unit[13] = (unsigned)B.hs * 4u;
    B.pf_ptr[14] = (const char*)B.face_lr;   B.pf_unit[14] = (unsigned)B.fs * 4u;
    B.pf_ptr[15] = (const char*)B.hdr;       B.pf_unit[15] = 32u;
  }
  for (int i = 0; i < 16; i++)   // the kernel's aligned prefetch path: rows and strides must be 16-byte granular
    if (B.pf_ptr[i] && (((uintptr_t)B.pf_ptr[i] & 15u) || (((size_t)EC * B.pf_unit[i]) & 15u && i < 13) ||
                        (i >= 13 && (B.pf_unit[i] & 15u))))
      B.pf_ptr[i] = nullptr;
  k<<<B.chunk_list ? P->n_generic : P->n_chunks, EC, smem, st>>>(B);
  return cudaGetLastError();
}

template <typename T, bool CMP>
static int launch_fused(const t8b200_plan* P, const FusedArgs<T>& A, cudaStream_t st) {
  // resident CTAs per SM are bounded by shared memory (68 KB fp64 / 36 KB fp32 per CTA); tell ptxas so it can size
  // the register budget (fp32: 5 CTAs of 48 registers measured faster than 6 of 40 and than 4 of 64)
#ifndef T8B_MINB64
#define T8B_MINB64 3
#endif
  constexpr int B0 = sizeof(T) == 8 ? T8B_MINB64 : 5;
  const bool split = P->split || A.chunk_list;   // a chunk list goes through the header-driven variant
  if (A.speed_max)
    return split ? launch_variant<T, MS, MF, B0, CMP, true, true>(P, A, st)
                 : launch_variant<T, MS, MF, B0, CMP, false, true>(P, A, st);
  return split ? launch_variant<T, MS, MF, B0, CMP, true, false>(P, A, st)
               : launch_variant<T, MS, MF, B0, CMP, false, false>(P, A, st);
}

template <typename T>
static int fused_stage_impl(const t8b200_plan* P, int stage, const T* const* in, const T* const* const* in_all,
                            const T* const* prev, T* const* out, const T* vol, T dt, T* speed_max, void* stream,
                            const T* dt_dev = nullptr, const t8b200_stage_sync* sync = nullptr,
                            long long wait_epoch = 0, long long signal_epoch = 0) {
  if (!P || stage < 1 || stage > 3 || !in || !out || !vol || (stage > 1 && !prev)) return cudaErrorInvalidValue;
  if (P->is_f64 != (sizeof(T) == 8) || P->host_only) return cudaErrorInvalidValue;
  if (P->multi && !P->ghost_tail && !in_all) return cudaErrorInvalidValue;   // (ghost-tail plans read no peer memory)
  if (P->n_chunks == 0) return cudaSuccess;   // a rank without elements: nothing to do (its rows may be null)
  // CTAs read halo states from `in` (or the peers' `in`) while other CTAs write `out`: an in-place call would race
  // silently; the reference-shaped rk3_stage tolerates out == prev, this entry point does not
  for (int k = 0; k < 5; k++)
    if (!in[k] || !out[k] || out[k] == in[k] || (stage > 1 && (!prev[k] || out[k] == prev[k]))) return cudaErrorInvalidValue;
  if (sync && (sync->nranks < 1 || sync->nranks > 32 || sync->rank < 0 || sync->rank >= sync->nranks ||
               !sync->mailboxes_dev || !sync->counter_dev || wait_epoch < 0 || signal_epoch < 0 || !P->multi ||
               P->ghost_tail))
    return cudaErrorInvalidValue;
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
  A.vol = vol; A.vol_shift = P->vol_shift; A.vol_scale = (T)P->vol_scale; A.dt = dt; A.speed_max = speed_max; A.n_local = P->n_local;
  A.stage = stage; A.multi = P->multi && !P->ghost_tail; A.my_rank = P->my_rank;
  A.dt_ptr = dt_dev;
  StageSync S{};
  if (sync) {
    S.mailboxes = (PeerSlot* const*)sync->mailboxes_dev; S.counter = sync->counter_dev;
    S.wait_epoch = wait_epoch; S.signal_epoch = signal_epoch;
    S.nranks = sync->nranks; S.rank = sync->rank; S.n_boundary_total = P->nb_struct + P->nb_generic;
    A.sync = S;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (speed_max) T8B_TRY(cudaMemsetAsync(speed_max, 0, sizeof(T), st));
  if (P->n_struct) {   // the structured chunks (structured.cu); what is left goes through the chunk list
    const int rc = t8b_structured_stage_run<T>(P, stage, in, in_all, prev, out, vol, dt, speed_max, stream, dt_dev,
                                               sync ? &S : nullptr);
    if (rc != 0 || P->n_generic == 0) return rc;
  }
  if (P->g_list) A.chunk_list = P->g_list;   // structured chunks elsewhere and / or boundary-first order (multi)
  return P->cmp ? launch_fused<T, true>(P, A, st) : launch_fused<T, false>(P, A, st);
}

// One pass of a stage split into "interior" (part 1: every chunk is launched, the partition-boundary ones leave at once)
// and "boundary" (part 2: the partition-boundary chunks only) launches, for ghost-tail plans whose chunks are all
// structured: part 1 needs no ghost copies, so it runs while the barrier and the pull of this stage proceed on another
// stream; part 2 follows the pull.  Both parts max into speed_max (cleared by the caller before both).
template <typename T>
static int fused_stage_part_impl(const t8b200_plan* P, int stage, int part, const T* const* in, const T* const* prev,
                                 T* const* out, const T* vol, T dt, const T* dt_dev, T* speed_max, void* stream) {
  if (!P || stage < 1 || stage > 3 || (part != 1 && part != 2) || !in || !out || !vol || (stage > 1 && !prev)) return cudaErrorInvalidValue;
  if (P->is_f64 != (sizeof(T) == 8) || P->host_only) return cudaErrorInvalidValue;
  if (!P->multi || !P->ghost_tail || P->n_struct != P->n_chunks || P->split || !P->blist) return cudaErrorNotSupported;
  for (int k = 0; k < 5; k++)
    if (!in[k] || !out[k] || out[k] == in[k] || (stage > 1 && (!prev[k] || out[k] == prev[k]))) return cudaErrorInvalidValue;
  // (speed_max is NOT zeroed here: the two passes run on different streams in either order; the caller clears it
  // before both)
  return t8b_structured_stage_run<T>(P, stage, in, nullptr, prev, out, vol, dt, speed_max, stream, dt_dev, nullptr, part);
}

// A stage of a ghost-tail plan whose chunks are all structured, with the push folded into the kernel.
template <typename T>
static int fused_stage_push_impl(const t8b200_plan* P, int stage, const T* const* in, const T* const* prev, T* const* out,
                                 T* const* const* out_all, const T* vol, T dt, const T* dt_dev, T* speed_max,
                                 const int32_t* send_off, const int32_t* send_rank, const int32_t* send_idx, void* stream) {
  if (!P || stage < 1 || stage > 3 || !in || !out || !out_all || !vol || (stage > 1 && !prev) || !send_off) return cudaErrorInvalidValue;
  if (P->is_f64 != (sizeof(T) == 8) || P->host_only) return cudaErrorInvalidValue;
  if (!P->multi || !P->ghost_tail || P->n_struct != P->n_chunks || P->split) return cudaErrorNotSupported;
  for (int k = 0; k < 5; k++)
    if (!in[k] || !out[k] || !out_all[k] || out[k] == in[k] || (stage > 1 && (!prev[k] || out[k] == prev[k]))) return cudaErrorInvalidValue;
  if (speed_max) T8B_TRY(cudaMemsetAsync(speed_max, 0, sizeof(T), (cudaStream_t)stream));
  t8b_push_args pa{};
  for (int k = 0; k < 5; k++) pa.out_all[k] = (const void* const*)out_all[k];
  pa.send_off = send_off; pa.send_rank = send_rank; pa.send_idx = send_idx;
  return t8b_structured_stage_run<T>(P, stage, in, nullptr, prev, out, vol, dt, speed_max, stream, dt_dev, nullptr, 0, &pa);
}

template <typename T>
int t8b_fused_stage_run(const t8b200_plan* P, int stage, const T* const* in, const T* const* const* in_all,
                        const T* const* prev, T* const* out, const T* vol, T dt, T* speed_max, void* stream,
                        const T* dt_dev, const t8b200_stage_sync* sync, long long wait_epoch, long long signal_epoch) {
  return fused_stage_impl<T>(P, stage, in, in_all, prev, out, vol, dt, speed_max, stream, dt_dev, sync, wait_epoch,
                             signal_epoch);
}
template int t8b_fused_stage_run<float>(const t8b200_plan*, int, const float* const*, const float* const* const*,
                                        const float* const*, float* const*, const float*, float, float*, void*,
                                        const float*, const t8b200_stage_sync*, long long, long long);
template int t8b_fused_stage_run<double>(const t8b200_plan*, int, const double* const*, const double* const* const*,
                                         const double* const*, double* const*, const double*, double, double*, void*,
                                         const double*, const t8b200_stage_sync*, long long, long long);

void t8b_plan_free(t8b200_plan* P) { t8b200_plan_destroy(P); }

// ============================================================================================================
// C ABI
// ============================================================================================================
extern "C" {

static int plan_create_impl(t8b200_plan** out, int host_only, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf,
                            int32_t nb, const int32_t* nbr, const void* normals, const void* areas,
                            const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                            const void* xnormals, const void* xareas) {
  if (!out || n_local < 0 || n_ghost < 0 || nf < 0 || nb < 0 || nx < 0) return cudaErrorInvalidValue;
  if ((nf + nb > 0) && (!nbr || !normals || !areas)) return cudaErrorInvalidValue;
  if (nx > 0 && (!xnbr || !xnormals || !xareas)) return cudaErrorInvalidValue;
  if (n_ghost > 0 && (!ranks || !indices)) return cudaErrorInvalidValue;
  t8b200_plan* P = new t8b200_plan();
  P->is_f64      = is_f64 ? 1 : 0;
  P->host_only   = host_only & 1;   // flags: bit 0 host-only plan, bit 1 ghost tail, bit 2 block program on the host
  P->ghost_tail  = (host_only >> 1) & 1;
  const bool emulate = (host_only & 5) == 5;   // plan_emulate.cuh: the device builder's per-block program, host loops
  int rc;
  if (is_f64) {
    MeshFaces<double> src{nf, nb, nx, nbr, (const double*)normals, (const double*)areas, ranks, indices, xnbr,
                          (const double*)xnormals, (const double*)xareas};
    rc = emulate ? plan_build_emulated<double>(P, n_local, n_ghost > 0, src) : plan_build<double>(P, n_local, n_ghost > 0, src);
  } else {
    MeshFaces<float> src{nf, nb, nx, nbr, (const float*)normals, (const float*)areas, ranks, indices, xnbr,
                         (const float*)xnormals, (const float*)xareas};
    rc = emulate ? plan_build_emulated<float>(P, n_local, n_ghost > 0, src) : plan_build<float>(P, n_local, n_ghost > 0, src);
  }
  if (rc != 0) {
    t8b200_plan_destroy(P);
    return rc;
  }
  *out = P;
  return 0;
}

int t8b200_plan_create(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb,
                       const int32_t* nbr, const void* normals, const void* areas, const int32_t* ranks,
                       const int32_t* indices, int32_t nx, const int32_t* xnbr, const void* xnormals,
                       const void* xareas) {
  return plan_create_impl(out, 0, is_f64, n_local, n_ghost, nf, nb, nbr, normals, areas, ranks, indices, nx, xnbr,
                          xnormals, xareas);
}
int t8b200_plan_create_host(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf, int32_t nb,
                            const int32_t* nbr, const void* normals, const void* areas, const int32_t* ranks,
                            const int32_t* indices, int32_t nx, const int32_t* xnbr, const void* xnormals,
                            const void* xareas) {
  return plan_create_impl(out, 1, is_f64, n_local, n_ghost, nf, nb, nbr, normals, areas, ranks, indices, nx, xnbr,
                          xnormals, xareas);
}

int t8b200_plan_create_block_program_host(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf,
                                          int32_t nb, const int32_t* nbr, const void* normals, const void* areas,
                                          const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                                          const void* xnormals, const void* xareas) {
  return plan_create_impl(out, 5, is_f64, n_local, n_ghost, nf, nb, nbr, normals, areas, ranks, indices, nx, xnbr,
                          xnormals, xareas);
}

int t8b200_plan_create_ghost_tail(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf,
                                  int32_t nb, const int32_t* nbr, const void* normals, const void* areas,
                                  const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                                  const void* xnormals, const void* xareas) {
  return plan_create_impl(out, 2, is_f64, n_local, n_ghost, nf, nb, nbr, normals, areas, ranks, indices, nx, xnbr,
                          xnormals, xareas);
}
int t8b200_plan_create_ghost_tail_host(t8b200_plan** out, int is_f64, int64_t n_local, int64_t n_ghost, int32_t nf,
                                       int32_t nb, const int32_t* nbr, const void* normals, const void* areas,
                                       const int32_t* ranks, const int32_t* indices, int32_t nx, const int32_t* xnbr,
                                       const void* xnormals, const void* xareas) {
  return plan_create_impl(out, 3, is_f64, n_local, n_ghost, nf, nb, nbr, normals, areas, ranks, indices, nx, xnbr,
                          xnormals, xareas);
}
int64_t t8b200_plan_ghost_tail_count(const t8b200_plan* P) { return P ? P->n_pull : -1; }

int t8b200_plan_host_array(const t8b200_plan* P, int which, const void** data, int64_t* count, int* elem_bytes) {
  if (!P || !P->host || !data || !count || !elem_bytes) return cudaErrorInvalidValue;
  const t8b200_plan_host& H = *P->host;
  auto set = [&](const auto& v) {
    *data       = v.data();
    *count      = (int64_t)v.size();
    *elem_bytes = (int)sizeof(v[0]);
    return 0;
  };
  switch (which) {
    case 0: return set(H.hdr);
    case 1: return set(H.halo_elem);
    case 2: return set(H.halo_rank);
    case 3: return set(H.face_lr);
    case 4: return set(H.face_ai);
    case 5: return set(H.ell);
    case 6: return set(H.ovf_off);
    case 7: return set(H.ovf_ent);
    case 8: return set(H.area_tab);
    case 9: return set(H.fnx);
    case 10: return set(H.fny);
    case 11: return set(H.fnz);
    case 12: return set(H.farea);
    case 13: return set(H.s_rec);
    case 14: return set(H.s_halo);
    case 15: return set(H.s_hrank);
    case 16: return set(H.g_list);
    case 17: return set(H.pull_rank);
    case 18: return set(H.pull_idx);
    case 19: return set(H.blist);
  }
  return cudaErrorInvalidValue;
}

void t8b200_plan_destroy(t8b200_plan* P) {
  if (!P) return;
  if (P->host_only) {
    delete P->host;
    delete P;
    return;
  }
  if (P->pool) {   // generic device builder: the chunk arrays are pieces of one allocation
    cudaFree(P->pool);
  } else {
    cudaFree(P->hdr); cudaFree(P->halo_elem); cudaFree(P->halo_rank);
    cudaFree(P->face_lr); cudaFree(P->face_ai); cudaFree(P->fnx); cudaFree(P->fny); cudaFree(P->fnz); cudaFree(P->farea);
    cudaFree(P->area_tab); cudaFree(P->ell); cudaFree(P->ovf_off); cudaFree(P->ovf_ent);
  }
  cudaFree(P->s_rec); cudaFree(P->s_halo); cudaFree(P->s_hrank); cudaFree(P->g_list);
  cudaFree(P->pull_rank); cudaFree(P->pull_idx); cudaFree(P->blist);
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
int t8b200_fused_stage_push_f32(const t8b200_plan* plan, int stage, const float* const* in, const float* const* prev,
                                float* const* out, float* const* const* out_all, const float* vol, float dt,
                                const float* dt_dev, float* speed_max_dev, const int32_t* send_off,
                                const int32_t* send_rank, const int32_t* send_idx, void* stream) {
  return fused_stage_push_impl<float>(plan, stage, in, prev, out, out_all, vol, dt, dt_dev, speed_max_dev, send_off,
                                      send_rank, send_idx, stream);
}
int t8b200_fused_stage_push_f64(const t8b200_plan* plan, int stage, const double* const* in, const double* const* prev,
                                double* const* out, double* const* const* out_all, const double* vol, double dt,
                                const double* dt_dev, double* speed_max_dev, const int32_t* send_off,
                                const int32_t* send_rank, const int32_t* send_idx, void* stream) {
  return fused_stage_push_impl<double>(plan, stage, in, prev, out, out_all, vol, dt, dt_dev, speed_max_dev, send_off,
                                       send_rank, send_idx, stream);
}
int t8b200_fused_stage_part_f32(const t8b200_plan* plan, int stage, int part, const float* const* in,
                                const float* const* prev, float* const* out, const float* vol, float dt,
                                const float* dt_dev, float* speed_max_dev, void* stream) {
  return fused_stage_part_impl<float>(plan, stage, part, in, prev, out, vol, dt, dt_dev, speed_max_dev, stream);
}
int t8b200_fused_stage_part_f64(const t8b200_plan* plan, int stage, int part, const double* const* in,
                                const double* const* prev, double* const* out, const double* vol, double dt,
                                const double* dt_dev, double* speed_max_dev, void* stream) {
  return fused_stage_part_impl<double>(plan, stage, part, in, prev, out, vol, dt, dt_dev, speed_max_dev, stream);
}
int t8b200_fused_stage_sync_f32(const t8b200_plan* plan, int stage, const float* const* in,
                                const float* const* const* in_all, const float* const* prev, float* const* out,
                                const float* vol, float dt, const float* dt_dev, float* speed_max_dev,
                                const t8b200_stage_sync* sync, long long wait_epoch, long long signal_epoch,
                                void* stream) {
  return fused_stage_impl<float>(plan, stage, in, in_all, prev, out, vol, dt, speed_max_dev, stream, dt_dev, sync,
                                 wait_epoch, signal_epoch);
}
int t8b200_fused_stage_sync_f64(const t8b200_plan* plan, int stage, const double* const* in,
                                const double* const* const* in_all, const double* const* prev, double* const* out,
                                const double* vol, double dt, const double* dt_dev, double* speed_max_dev,
                                const t8b200_stage_sync* sync, long long wait_epoch, long long signal_epoch,
                                void* stream) {
  return fused_stage_impl<double>(plan, stage, in, in_all, prev, out, vol, dt, speed_max_dev, stream, dt_dev, sync,
                                  wait_epoch, signal_epoch);
}

}  // extern "C"
