// The device tile-plan builder's per-block program (plan_block.cuh) driven by host loops: the same arrays through the
// same code the CUDA threads run, without a device -- the CPU-side check of device_plan.cu's generic path
// (t8b200_plan_create_host with flag bit 2; tests/test_plan_host_cpu.py compares every array with plan_build()).
// Host-only plans; single-rank plans get the structured / generic chunk lists, multi-rank ones stop at the chunk arrays
// (their launch order and ghost tail are device post-passes, checked on the GPU).
#pragma once
#include <algorithm>
#include <vector>

#include "plan_block.cuh"
#include "tile_plan.cuh"

template <typename T, typename Src>
static int plan_build_emulated(t8b200_plan* P, int64_t n_local, bool multi, Src& src) {
  namespace pb = t8b200::pb;
  static_assert(pb::EC == EC && pb::MS == MS && pb::MF == MF && pb::ELL == ELL, "plan_block.cuh constants");
  if (!P->host_only) return cudaErrorInvalidValue;
  const int64_t nblocks = (n_local + EC - 1) / EC, ntot = src.num_faces();
  P->n_local = n_local;
  P->multi   = multi ? 1 : 0;
  if (multi && n_local > 0) { int32_t rk = 0, ix = 0; src.owner(0, rk, ix); P->my_rank = rk; }
  // geometry classes
  bool           cmp = true;
  std::vector<T> area_tab;
  for (int64_t f = 0; f < ntot && cmp; f++) {
    T nrm[3], a;
    src.geometry(f, nrm, a);
    if (pb::axis_code_hd(nrm) < 0) cmp = false;
    else if (std::find(area_tab.begin(), area_tab.end(), a) == area_tab.end()) {
      area_tab.push_back(a);
      if (area_tab.size() > 256) cmp = false;
    }
  }
  std::sort(area_tab.begin(), area_tab.end());
  // faces by block, deliberately in DESCENDING id order: the program sorts its candidates
  std::vector<int64_t> face_off(nblocks + 1, 0);
  for (int64_t f = 0; f < ntot; f++) {
    int32_t l, r;
    src.endpoints(f, l, r);
    const int64_t cl = l < n_local ? l / EC : -1, cr = (r >= 0 && r < n_local) ? r / EC : -1;
    if (cl < 0 && cr < 0) return cudaErrorInvalidValue;
    if (cl >= 0) face_off[cl + 1]++;
    if (cr >= 0 && cr != cl) face_off[cr + 1]++;
  }
  for (int64_t b = 0; b < nblocks; b++) face_off[b + 1] += face_off[b];
  std::vector<int64_t> rec(face_off[nblocks]), cur(face_off.begin(), face_off.end() - 1);
  for (int64_t f = ntot - 1; f >= 0; f--) {
    int32_t l, r;
    src.endpoints(f, l, r);
    const int64_t cl = l < n_local ? l / EC : -1, cr = (r >= 0 && r < n_local) ? r / EC : -1;
    if (cl >= 0) rec[cur[cl]++] = f;
    if (cr >= 0 && cr != cl) rec[cur[cr]++] = f;
  }
  const int box_layout = P->vol_shift == 6 ? 1 : P->vol_shift == 0 ? 0 : -1;
  P->box_layout        = box_layout;
  int16_t inv[t8b200::SubgridBox::NSLOT];
  for (int h = 0; h < 256; h++)
    inv[box_layout == 1 ? t8b200::SubgridBox::thread_slot(h) : t8b200::MortonBox::thread_slot(h)] = (int16_t)h;
  int max_halo_allowed = MS - EC;
  if (const char* t = getenv("T8B200_TEST_MAX_HALO")) max_halo_allowed = std::min(max_halo_allowed, std::max(8, atoi(t)));
  pb::Params<T> pr{n_local, multi ? 1 : 0, cmp ? 1 : 0, (int)area_tab.size(), box_layout, max_halo_allowed, MF - 1,
                   area_tab.data(), inv};
  std::vector<unsigned char> arena((size_t)pb::Ws::bytes_per_program + 64);
  const pb::Ws               w = pb::Ws::carve(arena.data(), 1, 0, false);
  // COUNT
  std::vector<pb::Counts> cn(nblocks);
  pb::Out<T>              none{};
  for (int64_t b = 0; b < nblocks; b++) {
    pb::block_program<T, Src, false>(src, pr, w, b, rec.data() + face_off[b], face_off[b + 1] - face_off[b], cn[b], 0, 0,
                                     0, none);
    if (cn[b].rc) return cudaErrorInvalidValue;
  }
  std::vector<int64_t> cb(nblocks + 1, 0), ob(nblocks + 1, 0), eb(nblocks + 1, 0);
  int     max_halo = 0, max_faces = 0;
  int64_t n_halo = 0, n_rec = 0;
  bool    split = false;
  for (int64_t b = 0; b < nblocks; b++) {
    cb[b + 1] = cb[b] + cn[b].chunks; ob[b + 1] = ob[b] + cn[b].ovf_off; eb[b + 1] = eb[b] + cn[b].ovf_ent;
    max_halo = std::max(max_halo, cn[b].max_halo); max_faces = std::max(max_faces, cn[b].max_faces);
    n_halo += cn[b].sum_halo; n_rec += cn[b].sum_faces;
    split |= cn[b].chunks > 1;
  }
  const int64_t nchunks = cb[nblocks];
  if (n_local > 0x7FFFFF00LL || nchunks * MF > 0x7FFFFF00LL || ob[nblocks] > 0x7FFFFF00LL || eb[nblocks] > 0x7FFFFF00LL)
    return cudaErrorInvalidValue;
  const int HS = std::max(32, (max_halo + 31) / 32 * 32), FS = std::max(32, (max_faces + 31) / 32 * 32);
  auto* Hc = new t8b200_plan_host();
  P->host  = Hc;
  Hc->hdr.assign((size_t)nchunks * 8, 0);
  Hc->halo_elem.assign((size_t)nchunks * HS, -1);
  Hc->halo_rank.assign(multi ? (size_t)nchunks * HS : 0, 0);
  Hc->face_lr.assign((size_t)nchunks * FS, 0u);
  Hc->face_ai.assign(cmp ? (size_t)nchunks * FS : 0, (uint8_t)0);
  std::vector<T> fnx(cmp ? 0 : (size_t)nchunks * FS, T(0)), fny(fnx), fnz(fnx), far(fnx);
  Hc->ell.assign((size_t)std::max<int64_t>(n_local, 1) * ELL, (uint16_t)0xFFFF);
  Hc->ovf_off.assign((size_t)ob[nblocks], 0);
  Hc->ovf_ent.assign((size_t)eb[nblocks], 0);
  std::vector<uint8_t> s_flag((size_t)nchunks, 0);
  std::vector<int32_t> s_halo((size_t)nchunks * 256, 0), s_hrank(multi ? (size_t)nchunks * 256 : 0, 0);
  pb::Out<T> out{HS, FS, Hc->hdr.data(), Hc->halo_elem.data(), Hc->halo_rank.data(), Hc->face_lr.data(),
                 Hc->face_ai.data(), fnx.data(), fny.data(), fnz.data(), far.data(), Hc->ell.data(), Hc->ovf_off.data(),
                 Hc->ovf_ent.data(), s_flag.data(), s_halo.data(), s_hrank.data()};
  // FILL
  for (int64_t b = 0; b < nblocks; b++) {
    pb::Counts c2;
    pb::block_program<T, Src, true>(src, pr, w, b, rec.data() + face_off[b], face_off[b + 1] - face_off[b], c2, cb[b],
                                    ob[b], eb[b], out);
    if (c2.rc || c2.chunks != cn[b].chunks) return cudaErrorInvalidValue;
  }
  for (int64_t c = 0; c < nchunks; c++) {
    if (s_flag[c]) {
      Hc->s_rec.push_back(Hc->hdr[c * 8]); Hc->s_rec.push_back(Hc->hdr[c * 8 + 7]); Hc->s_rec.push_back((int32_t)c); Hc->s_rec.push_back(0);
      Hc->s_halo.insert(Hc->s_halo.end(), s_halo.begin() + c * 256, s_halo.begin() + (c + 1) * 256);
      if (multi) Hc->s_hrank.insert(Hc->s_hrank.end(), s_hrank.begin() + c * 256, s_hrank.begin() + (c + 1) * 256);
    } else Hc->g_list.push_back((int32_t)c);
  }
  P->n_chunks = (int)nchunks; P->split = split ? 1 : 0;
  P->n_struct = (int)(Hc->s_rec.size() / 4);
  P->n_generic = P->n_struct ? (int)Hc->g_list.size() : (int)nchunks;
  P->s_area0   = P->n_struct ? Hc->s_rec[1] : 0;
  if (!P->n_struct && !multi) Hc->g_list.clear();
  P->n_halo = n_halo; P->n_records = n_rec; P->hs = HS; P->fs = FS; P->max_halo = max_halo; P->max_faces = max_faces;
  P->ms = MS; P->mf = MF;
  P->smem_bytes = sizeof(T) == 8 ? 8 * ((size_t)t8b200::NCELLQ * MS + 5 * (size_t)MF) : 32 * (size_t)MS + 20 * (size_t)MF;
  P->cmp = cmp ? 1 : 0; P->n_areas = cmp ? (int)area_tab.size() : 0;
  if (cmp) Hc->area_tab.assign(area_tab.begin(), area_tab.end());
  Hc->fnx.assign(fnx.begin(), fnx.end()); Hc->fny.assign(fny.begin(), fny.end());
  Hc->fnz.assign(fnz.begin(), fnz.end()); Hc->farea.assign(far.begin(), far.end());
  return cudaSuccess;
}
