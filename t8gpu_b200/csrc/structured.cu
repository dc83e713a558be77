// Fused RK stage for the STRUCTURED chunks of a tile plan (box_layout.cuh): 256 consecutive elements / cells that form
// an 8 x 8 x 4 box with 256 single same-size face neighbours around it -- every chunk of a uniform hexahedral forest
// (BASELINE config 2), of a uniform Subgrid<4,4,4> forest (config 4), and the uniform regions of adaptive ones.  The
// generic kernel (fused.cu) keeps every other chunk (hanging faces, walls, general normals, 2-D, ragged ends).
//
// Same three phases as the generic kernel, but nothing is looked up:
//   phase 0  own conserved values (coalesced) and the 256 halo elements (gather through the per-chunk halo list, the
//            only plan data this kernel reads: 1 KB per chunk) -> 7 cell quantities in shared memory; the own cell stays
//            in registers
//   phase 1  every thread evaluates the three faces on the LOWER side of its element (left cell from shared memory,
//            right cell = its own registers), threads 0-127 one of the 128 faces on the upper box boundary
//   phase 2  out = RK combination of (sum of the three lower fluxes - sum of the three upper fluxes), indices arithmetic
// No face records, no element -> face table, no signs, no per-face branches; shared-memory accesses of the face phase
// are bank-conflict free by construction of the halo slots.
//
// Reference behaviour replaced (not translated): one stage of CompressibleEulerSolver::iterate
// (examples/compressible_euler/solver.cu:78-112) and of SubgridCompressibleEulerSolver::iterate
// (examples/subgrid/solver.inl:156-194) on the uniform parts of the mesh.
#if defined(T8B_S_PAIR) && T8B_S_PAIR
#define T8B_ENABLE_F32X2 1   // euler_flux.cuh: the packed two-face flux (experiment, off by default)
#endif
#include <algorithm>
#include <cstdlib>

#include "../../include/t8gpu_b200.h"
#include "box_layout.cuh"
#include "common.cuh"
#include "euler_flux.cuh"
#include "peer_sync.cuh"
#include "tile_plan.cuh"

using namespace t8b200;

template <typename T>
struct SArgs {
  const int4*    rec;     // per structured chunk: first element, area index, chunk id, -
  const int32_t* halo;    // 256 per chunk, thread order (box_layout.cuh), index into the owner's arrays
  const int32_t* hrank;   // owner rank (multi only)
  const T*       area_tab;
  const T*       in[5];
  const T* const* in_all[5];
  const T*       prev[5];
  T*             out[5];
  const T*       vol;
  int            vol_shift;
  T              vol_scale;
  T              dt;
  const T*       dt_ptr;  // non-null: the time step is read from device memory (t8b200_timestep_*)
  StageSync      sync;    // multi-GPU stage ordering done by the kernel itself (mailboxes == nullptr: off)
  T*             speed_max;
  int            stage, multi, my_rank;
  const int32_t* blist;   // MODE 3: ids of the partition-boundary chunks
  // MODE 4: the elements other ranks hold ghost copies of are pushed into those copies by the thread that writes them
  T* const*      out_all[5];   // [var][rank] tables of the OUTPUT step
  const int32_t *send_off, *send_rank, *send_idx;   // CSR by element: destinations (rank, index in that rank's rows)
  const uint4*   slots;   // T8B_S_TABLE: per thread (lower slots x | y << 16, z | ux << 16, uy | uz << 16, halo slot)
  int            dense;   // every chunk of the plan is structured: chunk b = elements [256 b, 256 b + 256), one area
  int            area0;
  int            wave;    // CTAs resident at once (distance of the next-wave L2 prefetch), 0: off
  int            pf_ok;   // rows are 16-byte aligned: bulk L2 prefetch hints allowed
};

__device__ __forceinline__ void s_prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <typename T, int NS>
__device__ __forceinline__ Cell<T> s_load_cell(const T* cq, int s) {
  Cell<T> q;
  q.rho = cq[0 * NS + s]; q.hx = cq[1 * NS + s]; q.hy = cq[2 * NS + s]; q.hz = cq[3 * NS + s];
  q.kp  = cq[4 * NS + s]; q.b  = cq[5 * NS + s]; q.q  = cq[6 * NS + s];
  return q;
}
template <typename T, int NS>
__device__ __forceinline__ void s_store_cell(T* cq, int s, const Cell<T>& q) {
  cq[0 * NS + s] = q.rho; cq[1 * NS + s] = q.hx; cq[2 * NS + s] = q.hy; cq[3 * NS + s] = q.hz;
  cq[4 * NS + s] = q.kp;  cq[5 * NS + s] = q.b;  cq[6 * NS + s] = q.q;
}

#ifndef T8B_S_MINB
#define T8B_S_MINB 3
#endif
#ifndef T8B_S_MINB32   // fp32: resident CTAs per SM the register budget is sized for.  6 (40 registers, 12 bytes of spills; 6 x
#define T8B_S_MINB32 6 // 34 KB of shared memory is all an SM holds) measured 1.428 ms per step against 1.463 at 5 x 48 registers
#endif
// 1: fp32 evaluates two faces per packed fp32x2 flux (fma.rn.f32x2, euler_flux.cuh: kepes_flux_x_pair).  Parity green,
// measured SLOWER on the level-8 hex forest (1.50 ms per step at 64 registers / 4 CTAs per SM, 1.57 at 48 / 5 with
// spills, 1.69 at 68 / 3, against 1.46 for the scalar flux at 48 / 5): the 210 packed instructions replace ~330 scalar
// ones, but packing / unpacking moves, the rotated cell copies and the lost occupancy cost more.  Off.
#ifndef T8B_S_PAIR
#define T8B_S_PAIR 0
#endif
#ifndef T8B_S_BULK     // 1: own elements through cp.async.bulk + mbarrier instead of LDG
#define T8B_S_BULK 0
#endif
// slot indices of a thread from a 256-entry table (one 16-byte load, L1 / L2 resident) instead of the layout's bit
// arithmetic: 0 off, 1 both precisions, 2 fp32 only.  fp32 is bound by issue slots and 40 % of its instructions are
// integer work: with the table 1.396 instead of 1.428 ms per step; fp64 loses 2.5 % (one more live register at the
// 80-register cap), so fp32 only.
#ifndef T8B_S_TABLE_SEL
#define T8B_S_TABLE_SEL 2
#endif
#define T8B_S_TABLE (T8B_S_TABLE_SEL == 1 || (T8B_S_TABLE_SEL == 2 && sizeof(T) == 4))
#ifndef T8B_S_OWNREG   // 1: the thread's own cell stays in registers through the face phase; 0: re-read per face
#define T8B_S_OWNREG 1
#endif

// MODE (its own instantiations, so that single-rank launches do not carry the others' registers: the kernel sits at the
// 80-register cap of 3 CTAs per SM):
//   0  every chunk of the launch, no peer memory (one rank, or a ghost-tail plan)
//   1  plans with ghosts read in place: peer tables, owner ranks, self-ordering against the peers
//   2  "interior pass" of a ghost-tail plan: every chunk is launched, the partition-boundary ones leave at once (their
//      ghost copies are still being pulled on another stream)
//   3  "boundary pass": the launch covers the partition-boundary chunks only, chunk id from the compact list
//   4  ghost-tail plan with the push folded in: a thread of a partition-boundary chunk also stores its new values into
//      the ghost copies the peers hold of its element (posted NVLink stores from the epilogue, no push kernel)
template <typename T, class L, bool SMAX, int MODE>
__global__ void __launch_bounds__(256, sizeof(T) == 8 ? T8B_S_MINB : T8B_S_MINB32)
    structured_stage_kernel(const __grid_constant__ SArgs<T> A) {
  constexpr bool MULTI = MODE == 1;
  constexpr int NS = L::NSLOT, NF = BoxCommon::NFLUX;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* const cq = reinterpret_cast<T*>(smem_raw);   // [7][NS]
  T* const fl = cq + NCELLQ * NS;                 // [5][NF]
  __shared__ T red[8];
  const int tid = threadIdx.x;
  const int b   = MODE == 3 ? __ldg(A.blist + blockIdx.x) : (int)blockIdx.x;

  // ---- phase 0: every independent global load first
  int  e0 = b * 256, area_idx = A.area0;
  bool bnd = false;   // partition-boundary chunk (reads ghost elements) of a launch that orders itself against the peers
  if (!A.dense) {
    const int4 r = __ldg(A.rec + b);
    e0 = r.x; area_idx = r.y;
    bnd = (MODE == 4 || (MULTI && A.sync.mailboxes != nullptr)) && r.w != 0;
  } else if ((MULTI && A.sync.mailboxes != nullptr) || MODE == 4) {
    bnd = __ldg(reinterpret_cast<const int*>(A.rec + b) + 3) != 0;   // off the critical path: only the flag is read
  }
  const int e    = e0 + tid;
  uint4     tab  = make_uint4(0u, 0u, 0u, 0u);
  if (T8B_S_TABLE) tab = __ldg(A.slots + tid);   // the thread's slot indices: the same for every chunk (L1 / L2 resident)
  const int hidx = __ldg(A.halo + b * 256 + tid);
  int       hrk  = A.my_rank;
  if (MULTI) hrk = __ldg(A.hrank + b * 256 + tid);
#if T8B_S_BULK
  // own elements: five bulk copies global -> shared (UBLKCP, completion on an mbarrier) issued by one thread into the
  // flux array, which is free until phase 1: no LSU issue slots and no landing registers for the 5 x 256 state values
  __shared__ __align__(8) unsigned long long mbar;
  const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(&mbar);
  const bool     bulk   = A.pf_ok != 0 && MODE != 2;   // (the interior pass may leave before the copies have landed)
  if (bulk) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      constexpr unsigned row_bytes = 256u * sizeof(T);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(5u * row_bytes) : "memory");
#pragma unroll
      for (int k = 0; k < 5; k++)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(fl + k * 256)),
                     "l"(A.in[k] + e0), "r"(row_bytes), "r"(mbar_a)
                     : "memory");
    }
  }
  T u0, u1, u2, u3, u4;
  if (!bulk) { u0 = A.in[0][e]; u1 = A.in[1][e]; u2 = A.in[2][e]; u3 = A.in[3][e]; u4 = A.in[4][e]; }
#else
  const T u0 = A.in[0][e], u1 = A.in[1][e], u2 = A.in[2][e], u3 = A.in[3][e], u4 = A.in[4][e];
#endif
  if (MODE == 2) {   // the flag was requested with everything else; a boundary chunk belongs to the other pass
    if (__ldg(reinterpret_cast<const int*>(A.rec + b) + 3) != 0) return;
  }
  if (A.pf_ok && (tid & 31) == 0) {
    // L2 prefetch hints, one or two per warp: the phase-2 operands of this chunk (rows of U^n, volume) and the streams
    // of the chunk that takes over this CTA slot about one wave later (CTAs are dispatched in index order)
    const int w = tid >> 5;
    if (w < 5) { if (A.stage != 1) s_prefetch_l2(A.prev[w] + e0, 256u * sizeof(T)); }
    else if (w == 5 && A.vol_shift == 0) s_prefetch_l2(A.vol + e0, 256u * sizeof(T));
    const int bw = b + A.wave;
    if (MODE != 3 && A.dense && A.wave > 0 && bw < (int)gridDim.x) {
      if (w < 5) s_prefetch_l2(A.in[w] + bw * 256, 256u * sizeof(T));
      else if (w == 5) s_prefetch_l2(A.halo + bw * 256, 1024u);
    }
  }
  T g0, g1, g2, g3, g4;
  if (MULTI && hrk != A.my_rank) {   // ghost: through the [var][rank] tables (a peer GPU's array over NVLink)
    // self-ordering launches: the owner's previous stage must be complete before its element is read
    if (A.sync.mailboxes != nullptr) stage_wait_owner(A.sync, hrk);
    g0 = A.in_all[0][hrk][hidx]; g1 = A.in_all[1][hrk][hidx]; g2 = A.in_all[2][hrk][hidx];
    g3 = A.in_all[3][hrk][hidx]; g4 = A.in_all[4][hrk][hidx];
  } else {
    g0 = A.in[0][hidx]; g1 = A.in[1][hidx]; g2 = A.in[2][hidx]; g3 = A.in[3][hidx]; g4 = A.in[4][hidx];
  }
#if T8B_S_BULK
  if (bulk) {
    __syncthreads();   // the barrier object is initialised before anybody waits on it
    unsigned done;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(mbar_a) : "memory");
    } while (!done);
    u0 = fl[tid]; u1 = fl[256 + tid]; u2 = fl[512 + tid]; u3 = fl[768 + tid]; u4 = fl[1024 + tid];
  }
#endif
  const Cell<T> C = to_cell(u0, u1, u2, u3, u4);
  s_store_cell<T, NS>(cq, tid, C);
  s_store_cell<T, NS>(cq, T8B_S_TABLE ? (int)tab.w : L::thread_slot(tid), to_cell(g0, g1, g2, g3, g4));
  __syncthreads();

  // ---- phase 1
  T smax = T(0);
#if T8B_S_PAIR
  if constexpr (sizeof(T) == 4) {
    // fp32: two faces per packed evaluation (kepes_flux_x_pair): (x, y) of this element, then z together with one of
    // the 128 faces on the upper box boundary (threads 0-127; the other threads duplicate their z face in that lane).
    // Each face is rotated into its own frame by a cyclic permutation of the velocity components.
    auto rot = [](const Cell<T>& c, int p0) {   // frame of a face with normal +e_p0: (h_p0, h_p0+1, h_p0+2)
      Cell<T> q = c;
      q.hx = p0 == 0 ? c.hx : p0 == 1 ? c.hy : c.hz;
      q.hy = p0 == 0 ? c.hy : p0 == 1 ? c.hz : c.hx;
      q.hz = p0 == 0 ? c.hz : p0 == 1 ? c.hx : c.hy;
      return q;
    };
    auto store = [&](int j, const T F[5], int p0) {   // momentum fluxes back to xyz
      fl[j] = F[0]; fl[4 * NF + j] = F[4];
      fl[(1 + p0) * NF + j] = F[1];
      fl[(1 + (p0 == 2 ? 0 : p0 + 1)) * NF + j] = F[2];
      fl[(1 + (p0 == 0 ? 2 : p0 - 1)) * NF + j] = F[3];
    };
    const int sl0 = L::at_lower(tid, 0) ? L::halo_slot(0, 0, L::compact(tid, 0)) : L::lower_own(tid, 0);
    const int sl1 = L::at_lower(tid, 1) ? L::halo_slot(1, 0, L::compact(tid, 1)) : L::lower_own(tid, 1);
    const int sl2 = L::at_lower(tid, 2) ? L::halo_slot(2, 0, L::compact(tid, 2)) : L::lower_own(tid, 2);
    T Fa[5], Fb[5], sa, sb;
    {
      const Cell<T> L0 = s_load_cell<T, NS>(cq, sl0), L1 = rot(s_load_cell<T, NS>(cq, sl1), 1), R1 = rot(C, 1);
      if (!kepes_flux_x_pair(L0, C, L1, R1, Fa, Fb, sa, sb)) {
        sa = kepes_flux_n<T, 0>(L0, C, T(0), T(0), T(0), Fa);
        sb = kepes_flux_n<T, 0>(L1, R1, T(0), T(0), T(0), Fb);
      }
      if (SMAX) smax = fmax_(smax, fmax_(sa, sb));
      store(tid, Fa, 0);
      store(256 + tid, Fb, 1);
    }
    {
      const bool up = tid < 128;
      const int  d = tid < 64 ? tid >> 5 : 2, idx = tid < 64 ? tid & 31 : tid - 64;
      const Cell<T> L0 = rot(s_load_cell<T, NS>(cq, sl2), 2), R0 = rot(C, 2);
      Cell<T>       L1 = L0, R1 = R0;
      if (up) {
        L1 = rot(s_load_cell<T, NS>(cq, L::upper_elem(idx, d)), d);
        R1 = rot(s_load_cell<T, NS>(cq, L::halo_slot(d, 1, idx)), d);
      }
      if (!kepes_flux_x_pair(L0, R0, L1, R1, Fa, Fb, sa, sb)) {
        sa = kepes_flux_n<T, 0>(L0, R0, T(0), T(0), T(0), Fa);
        sb = kepes_flux_n<T, 0>(L1, R1, T(0), T(0), T(0), Fb);
      }
      if (SMAX) smax = fmax_(smax, fmax_(sa, sb));
      store(512 + tid, Fa, 2);
      if (up) store(768 + tid, Fb, d);
    }
  } else
#endif
  {
#define T8B_S_FACE(D)                                                                                              \
  {                                                                                                                \
    const int sl = T8B_S_TABLE ? (D == 0 ? (int)(tab.x & 0xFFFFu) : D == 1 ? (int)(tab.x >> 16) : (int)(tab.y & 0xFFFFu)) \
                 : (L::at_lower(tid, D) ? L::halo_slot(D, 0, L::compact(tid, D)) : L::lower_own(tid, D));           \
    const Cell<T> Lc = s_load_cell<T, NS>(cq, sl);                                                                 \
    const Cell<T> Rc = T8B_S_OWNREG ? C : s_load_cell<T, NS>(cq, tid);                                             \
    T       F[5];                                                                                                  \
    const T s = kepes_flux_n<T, D>(Lc, Rc, T(0), T(0), T(0), F);                                                   \
    if (SMAX) smax = fmax_(smax, s);                                                                               \
    _Pragma("unroll") for (int k = 0; k < 5; k++) fl[k * NF + D * 256 + tid] = F[k];                               \
  }
  T8B_S_FACE(0)
  T8B_S_FACE(1)
  T8B_S_FACE(2)
#undef T8B_S_FACE
  if (tid < 128) {   // the 128 faces on the upper box boundary: warp 0 x+, warp 1 y+, warps 2-3 z+
    const int d = tid < 64 ? tid >> 5 : 2, idx = tid < 64 ? tid & 31 : tid - 64;
    const int p0 = d, p1 = d == 2 ? 0 : d + 1, p2 = d == 0 ? 2 : d - 1;   // cyclic axis permutation: one flux copy
    const int sl = L::upper_elem(idx, d), sr = L::halo_slot(d, 1, idx);
    Cell<T> Lc, Rc;
    Lc.rho = cq[sl]; Lc.hx = cq[(1 + p0) * NS + sl]; Lc.hy = cq[(1 + p1) * NS + sl]; Lc.hz = cq[(1 + p2) * NS + sl];
    Lc.kp = cq[4 * NS + sl]; Lc.b = cq[5 * NS + sl]; Lc.q = cq[6 * NS + sl];
    Rc.rho = cq[sr]; Rc.hx = cq[(1 + p0) * NS + sr]; Rc.hy = cq[(1 + p1) * NS + sr]; Rc.hz = cq[(1 + p2) * NS + sr];
    Rc.kp = cq[4 * NS + sr]; Rc.b = cq[5 * NS + sr]; Rc.q = cq[6 * NS + sr];
    T       F[5];
    const T s = kepes_flux_n<T, 0>(Lc, Rc, T(0), T(0), T(0), F);
    if (SMAX) smax = fmax_(smax, s);
    const int j = 768 + tid;   // == BoxCommon::upper_flux(d, idx)
    fl[j] = F[0]; fl[(1 + p0) * NF + j] = F[1]; fl[(1 + p1) * NF + j] = F[2]; fl[(1 + p2) * NF + j] = F[3];
    fl[4 * NF + j] = F[4];
  }

  }

  // ---- phase 2: operands requested before the barrier
  const int stage = A.stage;
  T         base[5];
  T         vol = A.vol[e >> A.vol_shift] * A.vol_scale;
#pragma unroll
  for (int k = 0; k < 5; k++) base[k] = A.in[k][e];
  if (stage != 1) {
    const T cp = stage == 2 ? T(0.75) : T(0.33333333333333), ci = stage == 2 ? T(0.25) : T(0.66666666666666);
#pragma unroll
    for (int k = 0; k < 5; k++) base[k] = cp * A.prev[k][e] + ci * base[k];
  }
  T sc = fast_rcp(vol) * (A.dt_ptr ? __ldg(A.dt_ptr) : A.dt);
  if (stage == 2) sc *= T(0.25);
  if (stage == 3) sc *= T(0.66666666666666);
  sc *= A.area_tab[area_idx];
  const int ux = T8B_S_TABLE ? (int)(tab.y >> 16)
                             : (L::at_upper(tid, 0) ? BoxCommon::upper_flux(0, L::compact(tid, 0)) : L::upper_own(tid, 0));
  const int uy = T8B_S_TABLE ? (int)(tab.z & 0xFFFFu)
                             : (L::at_upper(tid, 1) ? BoxCommon::upper_flux(1, L::compact(tid, 1)) : 256 + L::upper_own(tid, 1));
  const int uz = T8B_S_TABLE ? (int)(tab.z >> 16)
                             : (L::at_upper(tid, 2) ? BoxCommon::upper_flux(2, L::compact(tid, 2)) : 512 + L::upper_own(tid, 2));
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const T* f = fl + k * NF;
    const T  acc = ((f[tid] + f[256 + tid]) + f[512 + tid]) - ((f[ux] + f[uy]) + f[uz]);
    base[k] = base[k] + sc * acc;
    A.out[k][e] = base[k];
  }
  if (MODE == 4 && bnd) {   // the peers' copies of this element (0-3 of them), through the tables of the output step
    for (int q = __ldg(A.send_off + e), q1 = __ldg(A.send_off + e + 1); q < q1; q++) {
      const int rk = __ldg(A.send_rank + q), ix = __ldg(A.send_idx + q);
#pragma unroll
      for (int k = 0; k < 5; k++) A.out_all[k][rk][ix] = base[k];
    }
  }
  if (SMAX) {
    smax = warp_max(smax);
    if ((tid & 31) == 0) red[tid >> 5] = smax;
    __syncthreads();
    if (tid == 0) {
      T m = red[0];
#pragma unroll
      for (int w = 1; w < 8; w++) m = fmax_(m, red[w]);
      atomic_max_nonneg(A.speed_max, m);
    }
  }
  if (MULTI && bnd && A.sync.signal_epoch > 0) {   // every store of this chunk precedes the count (and the flag behind it)
    __syncthreads();
    if (tid == 0) stage_signal(A.sync);
  }
}

template <class L>
static const uint4* slot_table(int dev) {
  static uint4* tabs[64];
  if (!tabs[dev]) {
    uint4 h[256];
    for (int t = 0; t < 256; t++) {
      auto lower = [&](int d) { return L::at_lower(t, d) ? L::halo_slot(d, 0, L::compact(t, d)) : L::lower_own(t, d); };
      const unsigned ux = L::at_upper(t, 0) ? BoxCommon::upper_flux(0, L::compact(t, 0)) : L::upper_own(t, 0);
      const unsigned uy = L::at_upper(t, 1) ? BoxCommon::upper_flux(1, L::compact(t, 1)) : 256 + L::upper_own(t, 1);
      const unsigned uz = L::at_upper(t, 2) ? BoxCommon::upper_flux(2, L::compact(t, 2)) : 512 + L::upper_own(t, 2);
      h[t] = make_uint4((unsigned)lower(0) | (unsigned)lower(1) << 16, (unsigned)lower(2) | ux << 16, uy | uz << 16,
                        (unsigned)L::thread_slot(t));
    }
    if (cudaMalloc(&tabs[dev], sizeof(h)) != cudaSuccess) return nullptr;
    cudaMemcpy(tabs[dev], h, sizeof(h), cudaMemcpyHostToDevice);
  }
  return tabs[dev];
}

template <typename T, class L, bool SMAX, int MODE>
static int s_launch(const t8b200_plan* P, SArgs<T>& A, cudaStream_t st) {
  auto             k    = structured_stage_kernel<T, L, SMAX, MODE>;
  constexpr size_t smem = sizeof(T) * ((size_t)NCELLQ * L::NSLOT + 5 * (size_t)BoxCommon::NFLUX);
  // the opt-in above 48 KB is per device: cached per device id (not once per process)
  static int resident[64];
  int        dev = 0;
  T8B_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (resident[dev] == 0) {
    T8B_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0, per_sm = 0;
    T8B_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    T8B_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, 256, smem));
    resident[dev] = std::max(1, sms * per_sm);
  }
  static const int wave_knob = getenv("T8B200_WAVE") ? atoi(getenv("T8B200_WAVE")) : -1;
  A.wave = wave_knob >= 0 ? wave_knob : resident[dev];
  if (T8B_S_TABLE) {
    A.slots = slot_table<L>(dev);
    if (!A.slots) return cudaErrorMemoryAllocation;
  }
  const int grid = MODE == 3 ? P->nb_struct : P->n_struct;
  if (grid > 0) k<<<grid, 256, smem, st>>>(A);
  return cudaGetLastError();
}

template <typename T>
int t8b_structured_stage_run(const t8b200_plan* P, int stage, const T* const* in, const T* const* const* in_all,
                             const T* const* prev, T* const* out, const T* vol, T dt, T* speed_max, void* stream,
                             const T* dt_dev, const StageSync* sync, int part, const t8b_push_args* push) {
  if (P->n_struct == 0) return cudaSuccess;
  SArgs<T> A{};
  A.blist = P->blist;
  if (push) {
    for (int k = 0; k < 5; k++) A.out_all[k] = (T* const*)push->out_all[k];
    A.send_off = push->send_off; A.send_rank = push->send_rank; A.send_idx = push->send_idx;
  }
  A.dt_ptr = dt_dev;
  if (sync) A.sync = *sync;
  A.rec = reinterpret_cast<const int4*>(P->s_rec); A.halo = P->s_halo; A.hrank = P->s_hrank;
  A.area_tab = (const T*)P->area_tab;
  bool aligned = ((uintptr_t)vol & 15u) == 0 && ((uintptr_t)P->s_halo & 15u) == 0;
  for (int k = 0; k < 5; k++) {
    A.in[k]     = in[k];
    A.in_all[k] = in_all ? in_all[k] : nullptr;
    A.prev[k]   = stage > 1 ? prev[k] : in[k];
    A.out[k]    = out[k];
    aligned     = aligned && ((uintptr_t)A.in[k] & 15u) == 0 && ((uintptr_t)A.prev[k] & 15u) == 0;
  }
  A.vol = vol; A.vol_shift = P->vol_shift; A.vol_scale = (T)P->vol_scale; A.dt = dt; A.speed_max = speed_max;
  A.stage = stage; A.multi = P->multi && !P->ghost_tail; A.my_rank = P->my_rank;
  A.dense = (P->n_struct == P->n_chunks && !P->split) ? 1 : 0;   // (such plans keep their chunks in element order)
  static const bool no_dense = getenv("T8B200_TEST_NODENSE") != nullptr;   // timing experiment
  if (no_dense) A.dense = 0;
  A.area0 = P->s_area0;
  A.pf_ok = aligned ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  auto pick = [&](auto layout) {
    using L = decltype(layout);
    if (push) return speed_max ? s_launch<T, L, true, 4>(P, A, st) : s_launch<T, L, false, 4>(P, A, st);
    if (part == 1) return speed_max ? s_launch<T, L, true, 2>(P, A, st) : s_launch<T, L, false, 2>(P, A, st);
    if (part == 2) return speed_max ? s_launch<T, L, true, 3>(P, A, st) : s_launch<T, L, false, 3>(P, A, st);
    if (A.multi) return speed_max ? s_launch<T, L, true, 1>(P, A, st) : s_launch<T, L, false, 1>(P, A, st);
    return speed_max ? s_launch<T, L, true, 0>(P, A, st) : s_launch<T, L, false, 0>(P, A, st);
  };
  return P->box_layout == 1 ? pick(SubgridBox{}) : pick(MortonBox{});
}
template int t8b_structured_stage_run<float>(const t8b200_plan*, int, const float* const*, const float* const* const*,
                                             const float* const*, float* const*, const float*, float, float*, void*,
                                             const float*, const StageSync*, int, const t8b_push_args*);
template int t8b_structured_stage_run<double>(const t8b200_plan*, int, const double* const*,
                                              const double* const* const*, const double* const*, double* const*,
                                              const double*, double, double*, void*, const double*, const StageSync*,
                                              int, const t8b_push_args*);
