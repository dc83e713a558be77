// Tile plan shared by the unstructured (fused.cu) and the subgrid (subgrid.cu) fused stage: the connectivity of one
// rank re-laid out per chunk of EC = 256 consecutive elements (cells), built on the host from any face source.
//
// Plan layout (device), per chunk:
//   header   8 x int32: first element, elements, nh | nfc << 16, e0 | e1 << 16, e2, ovf_off_base, ovf_ent_base, area
//   halo     sorted unique elements outside the chunk that share a face with it (slot EC + h); fixed stride HS per
//            chunk (padding = -1), so the indices can be requested without waiting for the header
//   faces    one 32-bit record slotL (bits 0-13) | axis << 14 | slotR << 16 per face touching the chunk, fixed stride FS
//            per chunk (axis: compressed interior faces only, else 0).
//            Cartesian forests ("cmp": every normal +-e_axis, <= 256 distinct areas): records are put in canonical
//            orientation (normal = +e_axis, sides swapped where the stored normal was -e_axis) and grouped by axis,
//            [0,e0) x, [e0,e1) y, [e1,e2) z, then wall faces [e2,nfc) with the outward normal coded in the slotR field;
//            a chunk whose faces all have the same area carries it in the header (applied once per element).
//            General meshes: normals and areas as four T arrays, wall = slotR 0xFFFF.
//   ell      per element 8 x uint16 entries (face_local << 1 | side), 0xFFFF = none; one 128-bit load per thread.
//            Elements with more than 8 faces (hanging faces on several sides) continue in a per-chunk overflow CSR.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/t8gpu_b200.h"
#include "box_layout.cuh"
#include "euler_flux.cuh"

#ifndef T8B_EC
#define T8B_EC 256
#endif
static constexpr int EC  = T8B_EC;  // elements per chunk == threads per CTA
static constexpr int ELL = 8;    // face entries per element held in the fixed-width table

// std::vector that does not value-initialise on resize: the large per-chunk arrays of the merge are written in full by
// the parallel copy (payload + padding), a serial zero fill of hundreds of MB first would only add page-fault time
template <typename T>
struct default_init_allocator : std::allocator<T> {
  template <typename U> struct rebind { using other = default_init_allocator<U>; };
  template <typename U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
  template <typename U, typename... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
template <typename T>
using raw_vector = std::vector<T, default_init_allocator<T>>;

// host copy of the plan arrays (t8b200_plan_create_host: the builder without a device, for CPU-side checks)
struct t8b200_plan_host {
  raw_vector<int32_t>   hdr, halo_elem, halo_rank;
  raw_vector<uint32_t>  face_lr;
  raw_vector<uint8_t>   face_ai;
  raw_vector<uint16_t>  ell;
  std::vector<uint16_t> ovf_off, ovf_ent;
  std::vector<double>   area_tab, fnx, fny, fnz, farea;
  std::vector<int32_t>  s_rec, s_halo, s_hrank, g_list;   // structured chunks (box_layout.cuh)
  std::vector<int32_t>  pull_rank, pull_idx;              // ghost tail
  std::vector<int32_t>  blist;                            // partition-boundary structured chunks
};

struct t8b200_plan {
  int     is_f64     = 0;
  int64_t n_local    = 0;
  int     n_chunks   = 0;
  int     max_halo   = 0;
  int     max_faces  = 0;
  int     multi      = 0;  // has ghosts -> needs rank tables
  int     my_rank    = 0;  // owner rank of the local elements (halo entries of this rank skip the pointer tables)
  int     split      = 0;  // some blocks of EC elements were split: chunk c no longer starts at element c * EC
  size_t  smem_bytes = 0;
  int     ms = 0, mf = 0;  // compile-time stride variant selected for the kernel
  int64_t dev_bytes = 0, n_records = 0, n_halo = 0;
  int     cmp = 0, n_areas = 0;
  int     hs = 0, fs = 0;  // per-chunk strides of the halo and face arrays
  // device arrays (plans of the generic device builder: pieces of `pool`, one allocation)
  void*     pool      = nullptr;
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
  int64_t   n_ovf_off = 0, n_ovf_ent = 0;
  // structured chunks (box_layout.cuh, structured.cu): 8 x 8 x 4 boxes with 256 single same-size neighbours
  int      box_layout = -1;     // layout the builder tested the chunks against: 0 MortonBox, 1 SubgridBox, -1 none
  int      n_struct = 0, n_generic = 0;
  int      s_area0  = 0;        // area index shared by all chunks (dense plans)
  int32_t* s_rec    = nullptr;  // 4 per structured chunk: first element, area index, chunk id, 0
  int32_t* s_halo   = nullptr;  // 256 per structured chunk, thread order, index into the owner's arrays
  int32_t* s_hrank  = nullptr;  // owner ranks (multi only)
  int32_t* g_list   = nullptr;  // chunk ids left to the generic kernel (n_generic entries; when n_struct > 0 or multi)
  // multi: partition-boundary chunks (some halo element lives on another rank) come first in s_rec / s_halo / s_hrank
  // and in g_list, so that the stage kernels can wait for / signal the peers from those chunks alone (peer_sync.cuh)
  int      nb_struct = 0, nb_generic = 0;
  int32_t* blist = nullptr;     // ids of the partition-boundary structured chunks (nb_struct entries; dense plans)
  // volume lookup of the stage kernel: volume of element e = vol[e >> vol_shift] * vol_scale (subgrid cells share
  // their element's volume: shift 6 / 4, scale 1/64 / 1/16, ssp_runge_kutta.inl:116)
  int    vol_shift = 0;
  double vol_scale = 1.0;
  // ghost tail (t8b200_plan_create_ghost_tail): every ghost element a chunk reads has a LOCAL copy at index
  // n_local + j of this rank's own rows, filled before each stage by t8b200_ghost_pull_* from the peers' rows; the
  // halo entries point there (owner rank = this rank), so the stage kernels never touch peer memory
  int      ghost_tail = 0;
  int64_t  n_pull     = 0;
  int32_t* pull_rank  = nullptr;   // device, n_pull: owner rank of tail entry j
  int32_t* pull_idx   = nullptr;   // device, n_pull: its index in the owner's rows
  // host-only plans: arrays kept on the host, nothing uploaded, not launchable
  int               host_only = 0;
  t8b200_plan_host* host      = nullptr;
};


template <typename T, typename Alloc>
static T* upload(const std::vector<T, Alloc>& v, int64_t& bytes, cudaError_t& err) {
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

// Shared-memory geometry of the stage kernel: a chunk holds at most EC own elements, MS - EC halo elements and
// MF - 1 faces.  Chunks are EC consecutive elements; a block of EC elements whose halo or face count exceeds these
// limits (2:1 hanging faces, unstructured meshes with many small neighbours) is split recursively into smaller
// chunks, so ONE kernel variant serves every mesh.
#ifndef T8B_MS
#define T8B_MS 512
#define T8B_MF 1024
#endif
static constexpr int MS = T8B_MS, MF = T8B_MF;

// Host threads used to build a plan (per-block work is independent; results are merged in block order, so the plan
// does not depend on the thread count).
static inline int plan_threads() {
  if (const char* t = getenv("T8B200_PLAN_THREADS")) return std::max(1, atoi(t));
  unsigned hc = std::thread::hardware_concurrency();
  return (int)std::min(32u, std::max(1u, hc));
}
template <typename Fn>
static void parallel_ranges(int64_t n, int nthreads, Fn&& fn) {   // fn(thread, begin, end)
  nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, n));
  if (nthreads == 1) { fn(0, (int64_t)0, n); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back([&, t] { fn(t, n * t / nthreads, n * (t + 1) / nthreads); });
  for (auto& x : th) x.join();
}

// Launch order of a chunk list of a multi-rank plan (shared by the host builder below and the device builder,
// device_plan.cu).  flag[q] != 0: chunk q reads another rank's elements (partition-boundary chunk).  Boundary chunk j of
// nb sits at position floor(j W / nb) of the first W = max(nb, n/2) positions, interior chunks fill the rest in element
// order -- all boundary chunks are done (and signalled, peer_sync.cuh) about half way through the kernel, without a
// first wave made of nothing but NVLink-latency-bound chunks.  T8B200_BND_ORDER=first: all of them first; =natural:
// element order.  keep_order: plans whose chunks are all structured and unsplit keep the element order (chunk b =
// elements [256 b, 256 b + 256) needs no record load in front of the kernel's first loads: +4.5 % per step with it).
static inline int t8b_boundary_order_mode() {
  static const int mode = getenv("T8B200_BND_ORDER") ? (getenv("T8B200_BND_ORDER")[0] == 'f' ? 1 : getenv("T8B200_BND_ORDER")[0] == 'n' ? 2 : 0) : 0;
  return mode;
}
static inline std::vector<size_t> t8b_boundary_order(const std::vector<uint8_t>& flag, bool keep_order) {
  const int    order_mode = t8b_boundary_order_mode();
  const size_t n = flag.size();
  std::vector<size_t> bl, il, order;
  for (size_t q = 0; q < n; q++) (flag[q] ? bl : il).push_back(q);
  order.reserve(n);
  if (keep_order) { for (size_t q = 0; q < n; q++) order.push_back(q); return order; }
  if (order_mode == 1 || bl.empty()) { order = bl; order.insert(order.end(), il.begin(), il.end()); return order; }
  const size_t nb = bl.size(), W = std::max(nb, n / 2);
  size_t ib = 0, ii = 0;
  for (size_t pos = 0; pos < n; pos++) {
    const bool want_b = ib < nb && (pos >= W || ib * W / nb <= pos || ii >= il.size());
    if (want_b) order.push_back(bl[ib++]); else order.push_back(il[ii++]);
  }
  return order;
}

// Builds the plan from any face source `src` (its member functions are called concurrently, they must be const-safe):
//   int64_t num_faces();                                  faces of this rank, every face of a local element exactly once
//   void endpoints(int64_t f, int32_t& l, int32_t& r);    element ids; r = -1: wall; ids >= n_local: ghosts
//   void geometry(int64_t f, T nrm[3], T& area);          unit normal pointing l -> r (outward at a wall), face area
//   void owner(int32_t id, int32_t& rank, int32_t& idx);  owner rank and index in the owner's arrays (multi only)
template <typename T, typename Src>
static int plan_build(t8b200_plan* P, int64_t n_local, bool multi, Src& src) {
  const int nblocks = (int)((n_local + EC - 1) / EC);
  const int NT      = plan_threads();
  static const bool timing = getenv("T8B200_PLAN_TIMING") != nullptr;
  auto              t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[t8b200 plan] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  P->n_local = n_local;
  P->multi   = multi ? 1 : 0;
  if (multi && n_local > 0) {
    int32_t rk = 0, ix = 0;
    src.owner(0, rk, ix);
    P->my_rank = rk;
  }
  const int64_t ntot = src.num_faces();
  int max_halo_allowed = MS - EC;
  const int max_faces_allowed = MF - 1;
  if (const char* t = getenv("T8B200_TEST_MAX_HALO")) max_halo_allowed = std::min(max_halo_allowed, std::max(8, atoi(t)));
  static const bool by_id = getenv("T8B200_TEST_NOSORT") != nullptr;   // experiment: group order = face id order

  // ---- compressed geometry possible?  (every normal +-e_axis, <= 256 distinct areas): per-thread area sets, merged in
  // thread order, then one parallel pass assigns the table index of every face
  std::vector<std::vector<T>> local_areas(NT);
  std::vector<int>            not_axis(NT, 0);
  parallel_ranges(ntot, NT, [&](int t, int64_t f0, int64_t f1) {
    auto& la = local_areas[t];
    for (int64_t f = f0; f < f1; f++) {
      T nrm[3], a;
      src.geometry(f, nrm, a);
      if (axis_code(nrm) < 0) { not_axis[t] = 1; return; }
      if (!la.empty() && la.back() == a) continue;
      if (std::find(la.begin(), la.end(), a) == la.end()) {
        if (la.size() > 256) { not_axis[t] = 1; return; }
        la.push_back(a);
      }
    }
  });
  bool           cmp = true;
  std::vector<T> area_tab;
  for (int t = 0; t < NT && cmp; t++) {
    if (not_axis[t]) cmp = false;
    for (T a : local_areas[t])
      if (std::find(area_tab.begin(), area_tab.end(), a) == area_tab.end()) area_tab.push_back(a);
  }
  if (area_tab.size() > 256) cmp = false;
  std::sort(area_tab.begin(), area_tab.end());   // table order independent of the thread count
  raw_vector<uint8_t> area_of(cmp ? ntot : 0), code_of(cmp ? ntot : 0);   // per face: area index, axis_code of the normal
  if (cmp)
    parallel_ranges(ntot, NT, [&](int, int64_t f0, int64_t f1) {
      int last = 0;
      for (int64_t f = f0; f < f1; f++) {
        T nrm[3], a;
        src.geometry(f, nrm, a);
        if (area_tab[last] != a) last = (int)(std::find(area_tab.begin(), area_tab.end(), a) - area_tab.begin());
        area_of[f] = (uint8_t)last;
        code_of[f] = (uint8_t)axis_code(nrm);
      }
    });

  lap("geometry classes");
  // ---- bucket faces by block of EC elements (a face between two blocks appears in both): per-thread counts
  std::vector<std::vector<int32_t>> cnt(NT, std::vector<int32_t>(nblocks, 0));
  std::vector<int>                  bad(NT, 0);
  parallel_ranges(ntot, NT, [&](int t, int64_t f0, int64_t f1) {
    auto& c = cnt[t];
    for (int64_t f = f0; f < f1; f++) {
      int32_t l, r;
      src.endpoints(f, l, r);
      int cl = l < n_local ? l / EC : -1, cr = (r >= 0 && r < n_local) ? r / EC : -1;
      if (cl < 0 && cr < 0) { bad[t] = 1; return; }
      if (cl >= 0) c[cl]++;
      if (cr >= 0 && cr != cl) c[cr]++;
    }
  });
  for (int t = 0; t < NT; t++)
    if (bad[t]) return cudaErrorInvalidValue;
  std::vector<int64_t> face_off(nblocks + 1, 0);
  for (int b = 0; b < nblocks; b++) {
    int64_t sum = 0;
    for (int t = 0; t < NT; t++) { int32_t v = cnt[t][b]; cnt[t][b] = (int32_t)sum; sum += v; }   // thread offset in block
    face_off[b + 1] = face_off[b] + sum;
  }
  raw_vector<int64_t> rec(face_off[nblocks]);   // written in full by the pass below
  parallel_ranges(ntot, NT, [&](int t, int64_t f0, int64_t f1) {
    auto& c = cnt[t];
    for (int64_t f = f0; f < f1; f++) {
      int32_t l, r;
      src.endpoints(f, l, r);
      int cl = l < n_local ? l / EC : -1, cr = (r >= 0 && r < n_local) ? r / EC : -1;
      if (cl >= 0) rec[face_off[cl] + c[cl]++] = f;
      if (cr >= 0 && cr != cl) rec[face_off[cr] + c[cr]++] = f;
    }
  });
  cnt.clear();

  lap("bucket faces by block");
  // ---- per-block chunk emission, one builder per thread over a contiguous range of blocks
  raw_vector<uint16_t> ell((size_t)std::max<int64_t>(n_local, 1) * ELL);   // 0xFFFF-filled per chunk by emit()
  if (n_local == 0) std::fill(ell.begin(), ell.end(), (uint16_t)0xFFFF);
  struct Builder {
    std::vector<int32_t>  hdr, halo_elem, halo_rank;
    std::vector<int64_t>  halo_off{0}, rec_off{0};
    std::vector<uint32_t> face_lr;
    std::vector<uint8_t>  face_ai;
    std::vector<T>        fnx, fny, fnz, far;
    std::vector<uint16_t> ovf_off, ovf_ent;
    std::vector<uint8_t>  s_flag;             // per chunk: structured?
    std::vector<int32_t>  s_halo, s_hrank;    // 256 per structured chunk
    int  max_halo = 0, max_faces = 0, rc = 0;
    bool split = false;
  };
  // structured chunks: which box layout the chunks of this plan can have (unstructured plans: t8code hexahedra in
  // Morton order; cell-level plans of Subgrid<4,4,4>: 4 sibling elements); T8B200_STRUCTURED=0 leaves every chunk to the
  // generic kernel
  static const bool structured_on = !(getenv("T8B200_STRUCTURED") && atoi(getenv("T8B200_STRUCTURED")) == 0);
  const int box_layout = (structured_on && EC == 256 && MS - EC >= 256) ? (P->vol_shift == 6 ? 1 : P->vol_shift == 0 ? 0 : -1) : -1;
  P->box_layout = box_layout;
  int thread_of_slot[2][t8b200::SubgridBox::NSLOT];   // inverse of thread_slot() per layout
  for (int h = 0; h < 256; h++) {
    thread_of_slot[0][t8b200::MortonBox::thread_slot(h)]  = h;
    thread_of_slot[1][t8b200::SubgridBox::thread_slot(h)] = h;
  }
  std::vector<Builder> builders(NT);
  parallel_ranges(nblocks, NT, [&](int t, int64_t blk0, int64_t blk1) {
    Builder& B = builders[t];
    {   // sizes are known up to the splitting of blocks: no reallocation while the chunks are appended
      const size_t nrec = (size_t)(face_off[blk1] - face_off[blk0]), nblk = (size_t)(blk1 - blk0);
      B.face_lr.reserve(nrec);
      if (cmp) B.face_ai.reserve(nrec);
      else { B.fnx.reserve(nrec); B.fny.reserve(nrec); B.fnz.reserve(nrec); B.far.reserve(nrec); }
      B.hdr.reserve(8 * nblk); B.halo_elem.reserve((size_t)EC * nblk); B.halo_rank.reserve((size_t)EC * nblk);
      B.halo_off.reserve(nblk + 1); B.rec_off.reserve(nblk + 1);
    }
    std::vector<int32_t> halo_tmp;
    std::vector<int64_t> cand, sub;
    uint16_t              el_cnt[EC];
    int32_t               ht_key[1024];
    uint16_t              ht_val[1024];
    std::vector<uint32_t> ovf_pairs;
    struct Rec { int grp, sl, sr; int64_t f; };
    std::vector<Rec>      recs, recs_sorted;
    std::vector<int>      bucket;
    std::vector<int32_t>  ends;   // endpoints of the faces of the block being emitted (face sources compute them)

    // emits the chunk [b0, b1) whose faces are `faces`; 1 if it does not fit the kernel's shared memory
    auto emit = [&](int64_t b0, int64_t b1, const std::vector<int64_t>& faces, bool dry) -> int {
      const int nfc = (int)faces.size();
      halo_tmp.clear();
      ends.resize(2 * (size_t)nfc);
      for (int j = 0; j < nfc; j++) {
        int32_t l, r;
        src.endpoints(faces[j], l, r);
        ends[2 * j] = l; ends[2 * j + 1] = r;
        if (l < b0 || l >= b1) halo_tmp.push_back(l);
        if (r >= 0 && (r < b0 || r >= b1)) halo_tmp.push_back(r);
      }
      std::sort(halo_tmp.begin(), halo_tmp.end());
      halo_tmp.erase(std::unique(halo_tmp.begin(), halo_tmp.end()), halo_tmp.end());
      const int nh = (int)halo_tmp.size();
      if (nh > max_halo_allowed || nfc > max_faces_allowed) return 1;
      if (dry) return 0;
      B.max_halo  = std::max(B.max_halo, nh);
      B.max_faces = std::max(B.max_faces, nfc);
      B.hdr.resize(B.hdr.size() + 8, 0);
      int32_t* H = &B.hdr[B.hdr.size() - 8];
      H[0] = (int32_t)b0;
      H[1] = (int32_t)(b1 - b0);
      H[2] = nh | (nfc << 16);
      for (int h = 0; h < nh; h++) {
        int32_t id = halo_tmp[h], rk = 0, ix = id;
        if (multi) src.owner(id, rk, ix);
        else if (id >= n_local) return -1;   // a ghost without owner tables: invalid input (not "does not fit")
        B.halo_elem.push_back(ix);
        B.halo_rank.push_back(rk);
      }
      B.halo_off.push_back((int64_t)B.halo_elem.size());
      // halo id -> slot through a small open-addressing table (the list is sorted, but a probe beats the bisection)
      constexpr int HT = 1024;   // > 2 * (MS - EC) entries
      static_assert(HT >= 2 * (MS - EC), "hash table of the halo slots too small");
      for (int i = 0; i < HT; i++) ht_key[i] = -1;
      for (int h = 0; h < nh; h++) {
        unsigned k = ((unsigned)halo_tmp[h] * 2654435761u) >> 22;
        while (ht_key[k] >= 0) k = (k + 1) & (HT - 1);
        ht_key[k] = halo_tmp[h];
        ht_val[k] = (uint16_t)(EC + h);
      }
      auto slot_of = [&](int32_t id) -> int {
        if (id >= b0 && id < b1) return (int)(id - b0);
        unsigned k = ((unsigned)id * 2654435761u) >> 22;
        for (int probe = 0; ht_key[k] != id && probe < HT; probe++) k = (k + 1) & (HT - 1);   // always found
        return ht_val[k];
      };
      // kernel order of the records: cmp -> x, y, z interior faces, then walls; inside a group by left slot, so that
      // the threads of a warp read neighbouring slots of the cell array (few bank conflicts whatever the numbering)
      recs.resize(nfc);
      int seg[4] = {0, 0, 0, 0};
      for (int j = 0; j < nfc; j++) {
        const int64_t f = faces[j];
        const int32_t l = ends[2 * j], r = ends[2 * j + 1];
        int sl = slot_of(l), sr = r < 0 ? 0xFFFF : slot_of(r), grp = 0;
        if (cmp) {
          const int code = code_of[f];
          grp = r < 0 ? 3 : code >> 1;
          if (r < 0) sr = 0xFFF8 | code;                  // wall: outward normal coded in the slotR field
          else if (!(code & 1)) std::swap(sl, sr);        // canonical orientation: normal = +e_axis
          seg[grp]++;
        }
        recs[j] = Rec{grp, sl, sr, f};
      }
      if (by_id) {
        std::sort(recs.begin(), recs.end(), [](const Rec& x, const Rec& y) {
          if (x.grp != y.grp) return x.grp < y.grp;
          return x.f < y.f;
        });
      } else {
        // (group, left slot, right slot, face id) packed into one integer; the face id enters through the position in
        // `faces`, which is ascending in the id
        // counting sort on (group, left slot) -- stable, so the face id order survives -- then the few records that
        // share a left slot are ordered by right slot
        constexpr int NB = 4 * MS;
        bucket.assign(NB + 1, 0);
        for (int j = 0; j < nfc; j++) bucket[recs[j].grp * MS + std::min(recs[j].sl, MS - 1) + 1]++;
        for (int b = 0; b < NB; b++) bucket[b + 1] += bucket[b];
        std::vector<Rec>& sorted = recs_sorted;
        sorted.resize(nfc);
        for (int j = 0; j < nfc; j++) sorted[bucket[recs[j].grp * MS + std::min(recs[j].sl, MS - 1)]++] = recs[j];
        for (int j = 1; j < nfc; j++) {   // insertion sort inside runs of equal (group, left slot): (right slot, id)
          const Rec x = sorted[j];
          int       k = j;
          while (k > 0 && sorted[k - 1].grp == x.grp && sorted[k - 1].sl == x.sl &&
                 (sorted[k - 1].sr > x.sr || (sorted[k - 1].sr == x.sr && sorted[k - 1].f > x.f))) {
            sorted[k] = sorted[k - 1];
            k--;
          }
          sorted[k] = x;
        }
        recs.swap(sorted);
      }
      H[3] = seg[0] | ((seg[0] + seg[1]) << 16);
      H[4] = seg[0] + seg[1] + seg[2];
      // element -> face table written in place while the records are emitted (disjoint element ranges per chunk):
      // the first ELL entries of an element go to the fixed-width table, the rest to (slot, entry) pairs
      std::fill(ell.begin() + (size_t)b0 * ELL, ell.begin() + (size_t)b1 * ELL, (uint16_t)0xFFFF);
      std::fill(el_cnt, el_cnt + EC, (uint16_t)0);
      ovf_pairs.clear();
      auto add_entry = [&](int slot, uint16_t en) {
        const int k = el_cnt[slot]++;
        if (k < ELL) ell[((size_t)b0 + slot) * ELL + k] = en;
        else ovf_pairs.push_back(((uint32_t)slot << 16) | en);
      };
      int  area0 = -1;
      bool uniform = cmp;
      for (int j = 0; j < nfc; j++) {
        const int64_t f  = recs[j].f;
        const int     sl = recs[j].sl, sr = recs[j].sr;
        if (cmp) {
          B.face_ai.push_back(area_of[f]);
          if (area0 < 0) area0 = area_of[f];
          if (area_of[f] != area0) uniform = false;
        } else {
          T nrm[3], a;
          src.geometry(f, nrm, a);
          B.fnx.push_back(nrm[0]); B.fny.push_back(nrm[1]); B.fnz.push_back(nrm[2]); B.far.push_back(a);
        }
        // record: left slot (bits 0-13), axis of the normal (bits 14-15; compressed interior faces, else 0), right
        // slot or wall code (bits 16-31)
        const uint32_t axis_bits = (cmp && recs[j].grp < 3) ? (uint32_t)recs[j].grp << 14 : 0u;
        B.face_lr.push_back((uint32_t)sl | axis_bits | ((uint32_t)sr << 16));
        if (sl < EC) add_entry(sl, (uint16_t)(j << 1));
        if (sr < EC) add_entry(sr, (uint16_t)((j << 1) | 1));
      }
      B.rec_off.push_back((int64_t)B.face_lr.size());
      H[7] = (uniform && area0 >= 0) ? area0 : -1;
      // overflow CSR (bases relative to this builder): entries per element in emission order
      H[5] = -1;
      H[6] = 0;
      if (!ovf_pairs.empty()) {
        if (ovf_pairs.size() > 65535) return -1;
        H[5] = (int32_t)B.ovf_off.size();
        H[6] = (int32_t)B.ovf_ent.size();
        std::stable_sort(ovf_pairs.begin(), ovf_pairs.end(), [](uint32_t x, uint32_t y) { return (x >> 16) < (y >> 16); });
        size_t q = 0;
        for (int i = 0; i < EC; i++) {
          B.ovf_off.push_back((uint16_t)q);
          while (q < ovf_pairs.size() && (int)(ovf_pairs[q] >> 16) == i) B.ovf_ent.push_back((uint16_t)(ovf_pairs[q++] & 0xFFFFu));
        }
        B.ovf_off.push_back((uint16_t)q);
      }
      // structured?  full chunk, compressed geometry with one area, no walls, exactly 256 halo elements and 896 faces,
      // and the faces are those of the box: inside the box along the layout's index arithmetic, across its boundary to
      // ONE halo element per boundary element and direction (no hanging faces)
      bool structured = false;
      if (box_layout >= 0 && cmp && uniform && area0 >= 0 && b1 - b0 == 256 && (b0 & 255) == 0 && seg[3] == 0 &&
          nh == 256 && nfc == t8b200::BoxCommon::NFLUX) {
        int16_t lo[3][256], hi[3][256];
        std::fill(&lo[0][0], &lo[0][0] + 3 * 256, (int16_t)-1);
        std::fill(&hi[0][0], &hi[0][0] + 3 * 256, (int16_t)-1);
        structured = true;
        for (int j = 0; j < nfc && structured; j++) {   // canonical orientation: normal +e_grp from sl to sr
          const int d = recs[j].grp, sl = recs[j].sl, sr = recs[j].sr;
          if (sl < 256) { if (hi[d][sl] != -1) structured = false; hi[d][sl] = (int16_t)sr; }
          if (sr < 256) { if (lo[d][sr] != -1) structured = false; lo[d][sr] = (int16_t)sl; }
        }
        auto test = [&](auto tag) {
          using L = decltype(tag);
          const int* inv = thread_of_slot[box_layout];
          int32_t hl[256], hr[256];
          for (int t = 0; t < 256 && structured; t++)
            for (int d = 0; d < 3; d++) {
              const int l = lo[d][t], u = hi[d][t];
              if (L::at_lower(t, d) ? l < 256 : l != L::lower_own(t, d)) { structured = false; break; }
              if (L::at_upper(t, d) ? u < 256 : u != L::upper_own(t, d)) { structured = false; break; }
              if (L::at_lower(t, d)) {
                const int h = inv[L::halo_slot(d, 0, L::compact(t, d))], q = (int)B.halo_elem.size() - nh + (l - 256);
                hl[h] = B.halo_elem[q]; hr[h] = B.halo_rank[q];
              }
              if (L::at_upper(t, d)) {
                const int h = inv[L::halo_slot(d, 1, L::compact(t, d))], q = (int)B.halo_elem.size() - nh + (u - 256);
                hl[h] = B.halo_elem[q]; hr[h] = B.halo_rank[q];
              }
            }
          if (structured) {
            B.s_halo.insert(B.s_halo.end(), hl, hl + 256);
            B.s_hrank.insert(B.s_hrank.end(), hr, hr + 256);
          }
        };
        if (structured) { if (box_layout == 1) test(t8b200::SubgridBox{}); else test(t8b200::MortonBox{}); }
      }
      B.s_flag.push_back(structured ? 1 : 0);
      return 0;
    };
    // chunk [b0,b1) from the candidate faces of its block; halves it while it does not fit
    struct Range { int64_t b0, b1; };
    std::vector<Range> todo;
    for (int64_t blk = blk0; blk < blk1 && !B.rc; blk++) {
      cand.assign(rec.begin() + face_off[blk], rec.begin() + face_off[blk + 1]);
      const int64_t blk_b0 = blk * EC, blk_b1 = std::min<int64_t>(blk * EC + EC, n_local);
      todo.clear();
      todo.push_back({blk_b0, blk_b1});
      while (!todo.empty()) {
        const Range rg = todo.back();
        todo.pop_back();
        const std::vector<int64_t>* fs = &cand;
        if (rg.b1 - rg.b0 < blk_b1 - blk_b0) {   // part of a split block
          sub.clear();
          for (int64_t f : cand) {
            int32_t l, r;
            src.endpoints(f, l, r);
            if ((l >= rg.b0 && l < rg.b1) || (r >= rg.b0 && r < rg.b1)) sub.push_back(f);
          }
          fs = &sub;
        }
        // emit() checks the fit before it writes anything: no dry run needed
        int rc = emit(rg.b0, rg.b1, *fs, false);
        if (rc == 1) {
          if (rg.b1 - rg.b0 <= 1) { B.rc = cudaErrorInvalidValue; break; }   // one element exceeds a CTA
          const int64_t mid = (rg.b0 + rg.b1) / 2;
          todo.push_back({mid, rg.b1});   // LIFO: the lower half is emitted first, chunks stay in element order
          todo.push_back({rg.b0, mid});
          B.split = true;
          continue;
        }
        if (rc) { B.rc = rc < 0 ? (int)cudaErrorInvalidValue : rc; break; }
      }
    }
  });
  rec.clear();
  rec.shrink_to_fit();

  lap("chunk emission");
  // ---- merge the builders in block order, at fixed strides per chunk (halo padded with -1, faces with 0)
  int     nchunks = 0, max_halo = 0, max_faces = 0;
  bool    split = false;
  int64_t n_halo = 0, n_rec = 0;
  size_t  n_ovf_off = 0, n_ovf_ent = 0;
  for (auto& B : builders) {
    if (B.rc) return B.rc;
    nchunks += (int)(B.hdr.size() / 8);
    max_halo  = std::max(max_halo, B.max_halo);
    max_faces = std::max(max_faces, B.max_faces);
    split |= B.split;
    n_halo += (int64_t)B.halo_elem.size();
    n_rec += (int64_t)B.face_lr.size();
    n_ovf_off += B.ovf_off.size();
    n_ovf_ent += B.ovf_ent.size();
  }
  if (n_local > 0x7FFFFF00LL || (int64_t)nchunks * MF > 0x7FFFFF00LL || n_ovf_off > 0x7FFFFF00ULL ||
      n_ovf_ent > 0x7FFFFF00ULL)
    return cudaErrorInvalidValue;   // 32-bit indices in the kernel
  P->n_chunks = nchunks;
  P->split    = split ? 1 : 0;
  const int HS = std::max(32, (max_halo + 31) / 32 * 32), FS = std::max(32, (max_faces + 31) / 32 * 32);
  raw_vector<int32_t>  hdr((size_t)nchunks * 8), halo_elem((size_t)nchunks * HS),
      halo_rank(multi ? (size_t)nchunks * HS : 0);
  raw_vector<uint32_t> face_lr((size_t)nchunks * FS);
  raw_vector<uint8_t>  face_ai(cmp ? (size_t)nchunks * FS : 0);
  raw_vector<T>        fnx(cmp ? 0 : (size_t)nchunks * FS), fny(fnx.size()), fnz(fnx.size()), far(fnx.size());
  std::vector<uint16_t> ovf_off(n_ovf_off), ovf_ent(n_ovf_ent);
  struct BFlags { std::vector<uint8_t> flag; std::vector<int32_t> halo, hrank; };
  std::vector<BFlags> builders_flags(NT);
  for (int t = 0; t < NT; t++) {
    builders_flags[t].flag.swap(builders[t].s_flag);
    builders_flags[t].halo.swap(builders[t].s_halo);
    builders_flags[t].hrank.swap(builders[t].s_hrank);
  }
  {
    std::vector<int>    chunk_base(NT + 1, 0);
    std::vector<size_t> oo_base(NT + 1, 0), oe_base(NT + 1, 0);
    for (int t = 0; t < NT; t++) {
      chunk_base[t + 1] = chunk_base[t] + (int)(builders[t].hdr.size() / 8);
      oo_base[t + 1]    = oo_base[t] + builders[t].ovf_off.size();
      oe_base[t + 1]    = oe_base[t] + builders[t].ovf_ent.size();
    }
    parallel_ranges(NT, NT, [&](int, int64_t t0, int64_t t1) {
      for (int64_t t = t0; t < t1; t++) {
        Builder& B  = builders[t];
        const int nc = (int)(B.hdr.size() / 8);
        for (int i = 0; i < nc; i++) {
          const size_t c = (size_t)chunk_base[t] + i;
          int32_t*     H = &hdr[c * 8];
          std::copy(B.hdr.begin() + (size_t)i * 8, B.hdr.begin() + (size_t)i * 8 + 8, H);
          if (H[5] >= 0) { H[5] += (int32_t)oo_base[t]; H[6] += (int32_t)oe_base[t]; }
          // payload, then the padding up to the fixed stride (halo -1, everything else 0)
          auto place = [](auto& dst, size_t at, size_t stride, const auto& from, int64_t q0, int64_t q1, auto pad) {
            std::copy(from.begin() + q0, from.begin() + q1, dst.begin() + at);
            std::fill(dst.begin() + at + (size_t)(q1 - q0), dst.begin() + at + stride, pad);
          };
          place(halo_elem, c * HS, HS, B.halo_elem, B.halo_off[i], B.halo_off[i + 1], (int32_t)-1);
          if (multi) place(halo_rank, c * HS, HS, B.halo_rank, B.halo_off[i], B.halo_off[i + 1], (int32_t)0);
          const int64_t r0 = B.rec_off[i], r1 = B.rec_off[i + 1];
          place(face_lr, c * FS, FS, B.face_lr, r0, r1, (uint32_t)0);
          if (cmp) place(face_ai, c * FS, FS, B.face_ai, r0, r1, (uint8_t)0);
          else {
            place(fnx, c * FS, FS, B.fnx, r0, r1, T(0));
            place(fny, c * FS, FS, B.fny, r0, r1, T(0));
            place(fnz, c * FS, FS, B.fnz, r0, r1, T(0));
            place(far, c * FS, FS, B.far, r0, r1, T(0));
          }
        }
        std::copy(B.ovf_off.begin(), B.ovf_off.end(), ovf_off.begin() + oo_base[t]);
        std::copy(B.ovf_ent.begin(), B.ovf_ent.end(), ovf_ent.begin() + oe_base[t]);
        B = Builder();
      }
    });
  }
  // structured chunks: records and halo lists in chunk order, the other chunk ids for the generic kernel
  std::vector<int32_t> s_rec, s_halo, s_hrank, g_list, blist;
  {
    int c = 0;
    for (auto& B : builders_flags) {
      for (size_t i = 0, q = 0; i < B.flag.size(); i++, c++) {
        if (B.flag[i]) {
          s_rec.push_back(hdr[(size_t)c * 8]); s_rec.push_back(hdr[(size_t)c * 8 + 7]); s_rec.push_back(c); s_rec.push_back(0);
          s_halo.insert(s_halo.end(), B.halo.begin() + q * 256, B.halo.begin() + (q + 1) * 256);
          if (multi) s_hrank.insert(s_hrank.end(), B.hrank.begin() + q * 256, B.hrank.begin() + (q + 1) * 256);
          q++;
        } else g_list.push_back(c);
      }
    }
    P->n_struct  = (int)(s_rec.size() / 4);
    P->n_generic = P->n_struct ? (int)g_list.size() : nchunks;
    P->s_area0   = P->n_struct ? s_rec[1] : 0;
    if (!P->n_struct && !multi) g_list.clear();
    if (multi) {
      // partition-boundary chunks first (stable): the stage kernels order themselves against the peers from those
      // chunks alone (peer_sync.cuh); every rank must signal, so a rank without any gets one nominal boundary chunk
      const int32_t me = P->my_rank;
      std::vector<uint8_t> sb(P->n_struct, 0), gb(g_list.size(), 0);
      for (int q = 0; q < P->n_struct; q++)
        for (int i = 0; i < 256; i++)
          if (s_hrank[(size_t)q * 256 + i] != me) { sb[q] = 1; break; }
      for (size_t q = 0; q < g_list.size(); q++) {
        const size_t c = (size_t)g_list[q];
        for (int h = 0; h < HS; h++)
          if (halo_elem[c * HS + h] >= 0 && halo_rank[c * HS + h] != me) { gb[q] = 1; break; }
      }
      // order of the launches: t8b_boundary_order above.  The flag travels with the chunk: s_rec[4 q + 3] = 1, g_list
      // entry | 1 << 30.
      const bool keep_order = t8b_boundary_order_mode() == 2 || (P->n_struct == nchunks && !split);
      auto arrange = [&](const std::vector<uint8_t>& flag) { return t8b_boundary_order(flag, keep_order); };
      std::vector<int32_t> r2, h2, k2, g2;
      r2.reserve(s_rec.size()); h2.reserve(s_halo.size()); k2.reserve(s_hrank.size()); g2.reserve(g_list.size());
      for (size_t q : arrange(sb)) {
        r2.insert(r2.end(), s_rec.begin() + q * 4, s_rec.begin() + q * 4 + 4);
        r2.back() = sb[q];
        h2.insert(h2.end(), s_halo.begin() + q * 256, s_halo.begin() + q * 256 + 256);
        k2.insert(k2.end(), s_hrank.begin() + q * 256, s_hrank.begin() + q * 256 + 256);
        P->nb_struct += sb[q];
      }
      for (size_t q : arrange(gb)) {
        g2.push_back(g_list[q] | (gb[q] ? (1 << 30) : 0));
        P->nb_generic += gb[q];
      }
      s_rec.swap(r2); s_halo.swap(h2); s_hrank.swap(k2); g_list.swap(g2);
      if (P->nb_struct + P->nb_generic == 0 && nchunks > 0) {   // no ghosts at all: one nominal boundary chunk
        if (P->n_struct) { s_rec[3] = 1; P->nb_struct = 1; } else { g_list[0] |= 1 << 30; P->nb_generic = 1; }
      }
      for (int q = 0; q < P->n_struct; q++)
        if (s_rec[(size_t)q * 4 + 3]) blist.push_back(q);   // position in the structured launch (== chunk id when dense)
    }
  }
  // ghost tail: distinct (owner rank, remote index) pairs of the halo entries that live on other ranks, sorted ->
  // tail slot j; the entries are redirected to this rank's own rows at n_local + j.  The boundary flags above were
  // computed from the true owners and stay.
  std::vector<int32_t> pull_rank, pull_idx;
  if (P->ghost_tail && multi) {
    const int32_t me = P->my_rank;
    std::vector<uint64_t> keys;
    auto collect = [&](const auto& elem, const auto& rank) {
      for (size_t i = 0; i < elem.size(); i++)
        if (elem[i] >= 0 && rank[i] != me) keys.push_back(((uint64_t)(uint32_t)rank[i] << 32) | (uint32_t)elem[i]);
    };
    collect(halo_elem, halo_rank);
    collect(s_halo, s_hrank);
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    if ((int64_t)keys.size() + n_local > 0x7FFFFF00LL) return cudaErrorInvalidValue;
    auto redirect = [&](auto& elem, auto& rank) {
      for (size_t i = 0; i < elem.size(); i++)
        if (elem[i] >= 0 && rank[i] != me) {
          const uint64_t k = ((uint64_t)(uint32_t)rank[i] << 32) | (uint32_t)elem[i];
          elem[i] = (int32_t)(n_local + (std::lower_bound(keys.begin(), keys.end(), k) - keys.begin()));
          rank[i] = me;
        }
    };
    redirect(halo_elem, halo_rank);
    redirect(s_halo, s_hrank);
    pull_rank.resize(keys.size());
    pull_idx.resize(keys.size());
    for (size_t j = 0; j < keys.size(); j++) { pull_rank[j] = (int32_t)(keys[j] >> 32); pull_idx[j] = (int32_t)(keys[j] & 0xFFFFFFFFu); }
    P->n_pull = (int64_t)keys.size();
  }
  P->n_halo     = n_halo;
  P->n_records  = n_rec;
  P->n_ovf_off  = (int64_t)n_ovf_off;
  P->n_ovf_ent  = (int64_t)n_ovf_ent;
  P->hs = HS;
  P->fs = FS;
  P->max_halo   = max_halo;
  P->max_faces  = max_faces;
  P->ms = MS;
  P->mf = MF;
  // fp64: [7][MS] + [5][MF] doubles; fp32: two float4 per slot + float4 + float per face (fused.cu: Smem)
  P->smem_bytes = sizeof(T) == 8 ? 8 * ((size_t)t8b200::NCELLQ * MS + 5 * (size_t)MF) : 32 * (size_t)MS + 20 * (size_t)MF;

  lap("merge");
  P->cmp     = cmp ? 1 : 0;
  P->n_areas = cmp ? (int)area_tab.size() : 0;
  if (P->host_only) {
    auto* Hc = new t8b200_plan_host();
    Hc->hdr.swap(hdr); Hc->halo_elem.swap(halo_elem); Hc->halo_rank.swap(halo_rank); Hc->face_lr.swap(face_lr);
    Hc->face_ai.swap(face_ai); Hc->ell.swap(ell); Hc->ovf_off.swap(ovf_off); Hc->ovf_ent.swap(ovf_ent);
    if (cmp) Hc->area_tab.assign(area_tab.begin(), area_tab.end());
    Hc->fnx.assign(fnx.begin(), fnx.end()); Hc->fny.assign(fny.begin(), fny.end());
    Hc->fnz.assign(fnz.begin(), fnz.end()); Hc->farea.assign(far.begin(), far.end());
    Hc->s_rec.swap(s_rec); Hc->s_halo.swap(s_halo); Hc->s_hrank.swap(s_hrank); Hc->g_list.swap(g_list);
    Hc->pull_rank.swap(pull_rank); Hc->pull_idx.swap(pull_idx); Hc->blist.swap(blist);
    P->host = Hc;
    return cudaSuccess;
  }
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
  if (P->n_struct) {
    P->s_rec  = upload(s_rec, P->dev_bytes, err);
    P->s_halo = upload(s_halo, P->dev_bytes, err);
    if (P->multi) P->s_hrank = upload(s_hrank, P->dev_bytes, err);
  }
  if (!g_list.empty()) P->g_list = upload(g_list, P->dev_bytes, err);
  if (!blist.empty()) P->blist = upload(blist, P->dev_bytes, err);
  if (P->n_pull) {
    P->pull_rank = upload(pull_rank, P->dev_bytes, err);
    P->pull_idx  = upload(pull_idx, P->dev_bytes, err);
  }
  lap("upload");
  return err;
}

namespace t8b200 { struct StageSync; }
// push folded into the stage kernel (t8b200_fused_stage_push_*): tables of the output step + CSR of the destinations
struct t8b_push_args {
  const void* const* out_all[5];
  const int32_t *    send_off, *send_rank, *send_idx;
};
// defined in fused.cu (explicitly instantiated for float and double); dt_dev / sync: see t8b200_fused_stage_sync_*
template <typename T>
int t8b_fused_stage_run(const t8b200_plan* P, int stage, const T* const* in, const T* const* const* in_all,
                        const T* const* prev, T* const* out, const T* vol, T dt, T* speed_max, void* stream,
                        const T* dt_dev = nullptr, const t8b200_stage_sync* sync = nullptr, long long wait_epoch = 0,
                        long long signal_epoch = 0);
// defined in structured.cu: the structured chunks of the plan (no-op when it has none)
template <typename T>
int t8b_structured_stage_run(const t8b200_plan* P, int stage, const T* const* in, const T* const* const* in_all,
                             const T* const* prev, T* const* out, const T* vol, T dt, T* speed_max, void* stream,
                             const T* dt_dev, const t8b200::StageSync* sync, int part = 0,
                             const t8b_push_args* push = nullptr);
void t8b_plan_free(t8b200_plan* P);
// defined in subgrid.cu: wraps a cell-level plan (device_plan.cu builds them too)
t8b200_subgrid_plan* t8b_wrap_subgrid_plan(t8b200_plan* P, int dim);
