// The tile plan of ONE block of EC = 256 consecutive elements as a straight-line program without heap allocations or
// recursion, callable from a CUDA thread and from the host: what plan_build()'s per-block lambda (tile_plan.cuh) does
// with std::vector / std::sort, restated over a fixed workspace so that the device builder (device_plan.cu) can run one
// program per block with no host loop over the faces (SURVEY f-2).  The output arrays are the host builder's, bit for
// bit (tests/test_plan_host_cpu.py runs this program on the host through t8b200_plan_create_host, flag bit 2).
//
// Two passes: COUNT decides how the block splits into chunks and sizes them (chunks, overflow entries, maxima), FILL
// writes the chunks at the offsets the scan of those counts gives.
#pragma once
#include <cstdint>
#include <type_traits>

#include "box_layout.cuh"

namespace t8b200 {
namespace pb {

constexpr int EC = 256, MS = 512, MF = 1024, ELL = 8, HT = 1024;

// element i of the array of program `t`: base[i * stride + t] (device: interleaved, so that the threads of a warp that
// walk their arrays in step touch neighbouring words; host: stride 1)
template <typename U>
struct Arr {
  U*      p;
  int64_t stride;
  T8B_HD U& operator[](int64_t i) const { return p[i * stride]; }
};

struct Ws {
  Arr<int32_t>  ht_key;   // HT
  Arr<uint16_t> ht_val;   // HT
  Arr<uint16_t> el_cnt;   // EC
  Arr<uint16_t> el_run;   // EC
  Arr<int32_t>  halo;     // MS - EC
  Arr<uint64_t> key;      // MF
  Arr<int32_t>  pos;      // MF
  Arr<int32_t>  ends;     // 2 MF
  Arr<int16_t>  lo, hi;   // 3 EC each
  static constexpr int64_t bytes_per_program =
      4 * HT + 2 * HT + 2 * EC + 2 * EC + 4 * (MS - EC) + 8 * MF + 4 * MF + 8 * MF + 2 * (2 * 3 * EC);
  // carve the arrays of program t out of one arena for n programs (largest element type first: alignment)
  static T8B_HD Ws carve(unsigned char* base, int64_t n, int64_t t, bool interleave) {
    Ws            w;
    const int64_t s = interleave ? n : 1;
    auto take = [&](auto& arr, int64_t len) {
      using U = typename std::remove_reference<decltype(arr[0])>::type;
      arr.p      = reinterpret_cast<U*>(base) + (interleave ? t : t * len);
      arr.stride = s;
      base += sizeof(U) * len * n;
    };
    take(w.key, MF);
    take(w.ht_key, HT);
    take(w.halo, MS - EC);
    take(w.pos, MF);
    take(w.ends, 2 * MF);
    take(w.ht_val, HT);
    take(w.el_cnt, EC);
    take(w.el_run, EC);
    take(w.lo, 3 * EC);
    take(w.hi, 3 * EC);
    return w;
  }
};

template <typename T>
struct Params {
  int64_t        n_local;
  int            multi, cmp, n_areas, box_layout, max_halo_allowed, max_faces_allowed;
  const T*       area_tab;         // sorted ascending (cmp)
  const int16_t* thread_of_slot;   // inverse of L::thread_slot for the plan's box layout
};

template <typename T>
struct Out {   // FILL pass; per-chunk strides HS / FS
  int       HS, FS;
  int32_t * hdr, *halo_elem, *halo_rank;
  uint32_t* face_lr;
  uint8_t*  face_ai;
  T *       fnx, *fny, *fnz, *far;
  uint16_t *ell, *ovf_off, *ovf_ent;
  uint8_t*  s_flag;               // per chunk
  int32_t * s_halo, *s_hrank;     // 256 per chunk (valid where s_flag)
};

struct Counts {   // COUNT pass, per block
  int32_t chunks, ovf_off, ovf_ent, max_halo, max_faces, sum_halo, sum_faces, rc;
};

template <typename T>
__host__ __device__ inline int axis_code_hd(const T* n) {
  for (int d = 0; d < 3; d++) {
    const T o1 = n[(d + 1) % 3], o2 = n[(d + 2) % 3];
    if (o1 == T(0) && o2 == T(0) && (n[d] == T(1) || n[d] == T(-1))) return 2 * d + (n[d] > T(0) ? 1 : 0);
  }
  return -1;
}

// in-place heapsort (ascending); A indexable by [i]
template <class A, typename U>
__host__ __device__ inline void heapsort(const A& a, int64_t n, U) {
  auto sift = [&](int64_t root, int64_t end) {
    const U x = a[root];
    for (;;) {
      int64_t child = 2 * root + 1;
      if (child >= end) break;
      if (child + 1 < end && a[child] < a[child + 1]) child++;
      if (!(x < a[child])) break;
      a[root] = a[child];
      root    = child;
    }
    a[root] = x;
  };
  for (int64_t i = n / 2 - 1; i >= 0; i--) sift(i, n);
  for (int64_t end = n - 1; end > 0; end--) {
    const U t = a[0];
    a[0]      = a[end];
    a[end]    = t;
    sift(0, end);
  }
}

__host__ __device__ inline unsigned halo_hash(int32_t id) { return ((unsigned)id * 2654435761u) >> 22; }

// One chunk [b0, b1) out of the candidate faces of its block.  Returns 1 when it does not fit the kernel's shared
// memory (the caller halves it), < 0 on invalid input, 0 when counted / written.
template <typename T, class Src, bool FILL>
__host__ __device__ inline int chunk_program(const Src& src, const Params<T>& pr, const Ws& w, const int64_t* cand, int64_t ncand,
                                bool whole, int64_t b0, int64_t b1, Counts& cn, int64_t c, int64_t oo_at,
                                int64_t oe_at, const Out<T>& out) {
  // the faces of the range, ascending in the face id
  int nfc = 0;
  for (int64_t q = 0; q < ncand; q++) {
    int32_t l, r;
    src.endpoints(cand[q], l, r);
    if (!whole && !((l >= b0 && l < b1) || (r >= b0 && r < b1))) continue;
    if (nfc >= pr.max_faces_allowed) return 1;
    w.pos[nfc]          = (int32_t)q;
    w.ends[2 * nfc]     = l;
    w.ends[2 * nfc + 1] = r;
    nfc++;
  }
  // distinct elements outside the range that share a face with it
  for (int i = 0; i < HT; i++) w.ht_key[i] = -1;
  int nh = 0;
  for (int j = 0; j < 2 * nfc; j++) {
    const int32_t id = w.ends[j];
    if (id < 0 || (id >= b0 && id < b1)) continue;
    unsigned k = halo_hash(id);
    while (w.ht_key[k] >= 0 && w.ht_key[k] != id) k = (k + 1) & (HT - 1);
    if (w.ht_key[k] < 0) {
      if (nh >= pr.max_halo_allowed) return 1;
      w.ht_key[k] = id;
      w.halo[nh++] = id;
    }
  }
  // entries of the element -> face table per own slot; beyond ELL they go to the overflow CSR
  for (int i = 0; i < EC; i++) w.el_cnt[i] = 0;
  for (int j = 0; j < 2 * nfc; j++) {
    const int32_t id = w.ends[j];
    if (id >= b0 && id < b1) w.el_cnt[id - b0] = (uint16_t)(w.el_cnt[id - b0] + 1);
  }
  int n_ovf = 0;
  for (int i = 0; i < EC; i++) n_ovf += w.el_cnt[i] > ELL ? w.el_cnt[i] - ELL : 0;
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
  H[0] = (int32_t)b0;
  H[1] = (int32_t)(b1 - b0);
  H[2] = nh | (nfc << 16);
  heapsort(w.halo, nh, int32_t{});
  for (int h = 0; h < nh; h++) {
    const int32_t id = w.halo[h];
    int32_t       rk = 0, ix = id;
    if (pr.multi) src.owner(id, rk, ix);
    else if (id >= pr.n_local) return -1;   // a ghost without owner tables
    out.halo_elem[c * out.HS + h] = ix;
    if (pr.multi) out.halo_rank[c * out.HS + h] = rk;
    unsigned k = halo_hash(id);
    while (w.ht_key[k] != id) k = (k + 1) & (HT - 1);
    w.ht_val[k] = (uint16_t)(EC + h);
  }
  auto slot_of = [&](int32_t id) -> int {
    if (id >= b0 && id < b1) return (int)(id - b0);
    unsigned k = halo_hash(id);
    while (w.ht_key[k] != id) k = (k + 1) & (HT - 1);
    return w.ht_val[k];
  };
  // kernel order of the records: (group, left slot, right slot, face id) as one integer
  int seg[4] = {0, 0, 0, 0};
  for (int j = 0; j < nfc; j++) {
    const int32_t l = w.ends[2 * j], r = w.ends[2 * j + 1];
    int           sl = slot_of(l), sr = r < 0 ? 0xFFFF : slot_of(r), grp = 0;
    if (pr.cmp) {
      T nrm[3], a;
      src.geometry(cand[w.pos[j]], nrm, a);
      const int code = axis_code_hd(nrm);
      grp = r < 0 ? 3 : code >> 1;
      if (r < 0) sr = 0xFFF8 | code;
      else if (!(code & 1)) { const int t = sl; sl = sr; sr = t; }
      seg[grp]++;
    }
    w.key[j] = ((uint64_t)(grp * MS + sl) << 32) | ((uint64_t)sr << 16) | (uint64_t)j;
  }
  heapsort(w.key, nfc, uint64_t{});
  H[3] = seg[0] | ((seg[0] + seg[1]) << 16);
  H[4] = seg[0] + seg[1] + seg[2];
  H[5] = -1;
  H[6] = 0;
  if (n_ovf) {
    H[5] = (int32_t)oo_at;
    H[6] = (int32_t)oe_at;
    int q = 0;
    for (int i = 0; i < EC; i++) {
      out.ovf_off[oo_at + i] = (uint16_t)q;
      q += w.el_cnt[i] > ELL ? w.el_cnt[i] - ELL : 0;
    }
    out.ovf_off[oo_at + EC] = (uint16_t)q;
  }
  for (int i = 0; i < EC; i++) w.el_run[i] = 0;
  auto add_entry = [&](int slot, uint16_t en) {
    const int k   = w.el_run[slot];
    w.el_run[slot] = (uint16_t)(k + 1);
    if (k < ELL) out.ell[(b0 + slot) * ELL + k] = en;
    else out.ovf_ent[oe_at + out.ovf_off[oo_at + slot] + (k - ELL)] = en;
  };
  int  area0   = -1;
  bool uniform = pr.cmp != 0;
  for (int jj = 0; jj < nfc; jj++) {
    const uint64_t kk = w.key[jj];
    const int      j = (int)(kk & 0xFFFFu), sr = (int)((kk >> 16) & 0xFFFFu), gs = (int)(kk >> 32), grp = gs / MS, sl = gs % MS;
    T              nrm[3], a;
    src.geometry(cand[w.pos[j]], nrm, a);
    if (pr.cmp) {
      int lo = 0, hi = pr.n_areas - 1;   // exact match exists
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (pr.area_tab[mid] < a) lo = mid + 1; else hi = mid; }
      out.face_ai[c * out.FS + jj] = (uint8_t)lo;
      if (area0 < 0) area0 = lo;
      if (lo != area0) uniform = false;
    } else {
      out.fnx[c * out.FS + jj] = nrm[0]; out.fny[c * out.FS + jj] = nrm[1]; out.fnz[c * out.FS + jj] = nrm[2];
      out.far[c * out.FS + jj] = a;
    }
    const uint32_t axis_bits = (pr.cmp && grp < 3) ? (uint32_t)grp << 14 : 0u;
    out.face_lr[c * out.FS + jj] = (uint32_t)sl | axis_bits | ((uint32_t)sr << 16);
    if (sl < EC) add_entry(sl, (uint16_t)(jj << 1));
    if (sr < EC) add_entry(sr, (uint16_t)((jj << 1) | 1));
    w.key[jj] = ((uint64_t)grp << 32) | ((uint64_t)sr << 16) | (uint64_t)sl;   // kept for the structured test below
  }
  H[7] = (uniform && area0 >= 0) ? area0 : -1;

  // structured?  (tile_plan.cuh: full aligned chunk, one area, no walls, 256 halo elements, the faces of the box)
  bool structured = false;
  if (pr.box_layout >= 0 && pr.cmp && uniform && area0 >= 0 && b1 - b0 == 256 && (b0 & 255) == 0 && seg[3] == 0 &&
      nh == 256 && nfc == BoxCommon::NFLUX) {
    structured = true;
    for (int i = 0; i < 3 * EC; i++) { w.lo[i] = -1; w.hi[i] = -1; }
    for (int jj = 0; jj < nfc && structured; jj++) {
      const uint64_t kk = w.key[jj];
      const int      d = (int)(kk >> 32), sr = (int)((kk >> 16) & 0xFFFFu), sl = (int)(kk & 0xFFFFu);
      if (sl < 256) { if (w.hi[d * EC + sl] != -1) structured = false; w.hi[d * EC + sl] = (int16_t)sr; }
      if (sr < 256) { if (w.lo[d * EC + sr] != -1) structured = false; w.lo[d * EC + sr] = (int16_t)sl; }
    }
    auto test = [&](auto tag) {
      using L = decltype(tag);
      for (int t = 0; t < 256 && structured; t++)
        for (int d = 0; d < 3; d++) {
          const int l = w.lo[d * EC + t], u = w.hi[d * EC + t];
          if (L::at_lower(t, d) ? l < 256 : l != L::lower_own(t, d)) { structured = false; break; }
          if (L::at_upper(t, d) ? u < 256 : u != L::upper_own(t, d)) { structured = false; break; }
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
    if (pr.box_layout == 1) test(SubgridBox{}); else test(MortonBox{});
  }
  out.s_flag[c] = structured ? 1 : 0;
  cn.chunks++;
  if (n_ovf) { cn.ovf_off += EC + 1; cn.ovf_ent += n_ovf; }
  return 0;
}

// The block `blk`: halves the range while a chunk does not fit (lower half first: chunks stay in element order).
// COUNT: cand is sorted in place and `cn` receives the block's counts.  FILL: chunk_base / oo_base / oe_base are the
// exclusive scans of the counts.
template <typename T, class Src, bool FILL>
__host__ __device__ inline void block_program(const Src& src, const Params<T>& pr, const Ws& w, int64_t blk, int64_t* cand,
                                 int64_t ncand, Counts& cn, int64_t chunk_base, int64_t oo_base, int64_t oe_base,
                                 const Out<T>& out) {
  struct Range { int64_t b0, b1; };
  Range         todo[40];
  int           sp = 0;
  const int64_t blk_b0 = blk * EC, blk_b1 = blk * EC + EC < pr.n_local ? blk * EC + EC : pr.n_local;
  cn = Counts{0, 0, 0, 0, 0, 0, 0, 0};
  if (!FILL) heapsort(cand, ncand, int64_t{});
  todo[sp++] = {blk_b0, blk_b1};
  while (sp > 0) {
    const Range rg = todo[--sp];
    const int   rc = chunk_program<T, Src, FILL>(src, pr, w, cand, ncand, rg.b1 - rg.b0 == blk_b1 - blk_b0, rg.b0, rg.b1, cn,
                                                 chunk_base + cn.chunks, oo_base + cn.ovf_off, oe_base + cn.ovf_ent, out);
    if (rc == 1) {
      if (rg.b1 - rg.b0 <= 1) { cn.rc = 1; return; }   // one element exceeds a CTA
      const int64_t mid = (rg.b0 + rg.b1) / 2;
      todo[sp].b0 = mid;   todo[sp++].b1 = rg.b1;
      todo[sp].b0 = rg.b0; todo[sp++].b1 = mid;
      continue;
    }
    if (rc) { cn.rc = 1; return; }
  }
}

}  // namespace pb
}  // namespace t8b200
