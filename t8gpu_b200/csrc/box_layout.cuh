// Index arithmetic of a "structured chunk": 256 consecutive elements (cells) that form an 8 x 8 x 4 box whose 256 face
// neighbours outside the box are single same-size elements.  For such a chunk the stage kernel (structured.cu) needs no
// face records and no element -> face table: every index below is arithmetic on the thread id.
//
// Two orderings of the 256 elements occur on the path:
//   MortonBox   t8code hexahedra of one level in SFC order, bits of t = x0 y0 z0 x1 y1 z1 x2 y2 (x fastest)
//               (uniform regions of a MeshManager forest; the element order t8_forest_new_uniform / adapt produce)
//   SubgridBox  the cells of 4 consecutive Subgrid<4,4,4> elements that are Morton siblings in x and y:
//               t = el * 64 + i + 4 j + 16 k, el = ex + 2 ey  (t8gpu/memory/subgrid_memory_manager.h:35-68)
//
// Shared-memory slots: own element t -> slot t; the 256 halo elements -> slots >= 256, placed so that a half-warp that
// reads "my lower neighbour along d" (own slots for most lanes, halo slots for the lanes on the box boundary) touches
// 16 different 8-byte banks: the halo slot of a boundary lane gets the bank its lower neighbour would have had if the
// box continued.  That costs holes in the halo region (NSLOT > 512) and removes the bank conflicts of the generic
// plan's sorted halo list (18 % shared-memory replays, profiles/r1_fused_f64_ncu_summary.csv).
//
// Halo list order: entry h in [0,256) is staged by thread h into slot thread_slot(h) = the h-th used halo slot in
// ascending slot order (so that the stores of a half-warp spread over the banks); halo_slot(d, side, idx) names the slot
// of the neighbour across the lower (side 0) / upper (side 1) box face along d of the boundary element with compact
// index idx in that face (compact()).
// Flux slots: lower face of element t along d at d * 256 + t; the faces on the upper box boundary at 768 + [x+ 32]
// [y+ 32][z+ 64] by compact index.
#pragma once
#include "euler_flux.cuh"

namespace t8b200 {

struct MortonBox {
  static constexpr int NSLOT = 256 + 320;
  T8B_HD static constexpr unsigned mask(int d) { return d == 0 ? 0x49u : d == 1 ? 0x92u : 0x24u; }   // bits of t holding coordinate d
  T8B_HD static int coord(int t, int d) {
    return d == 0 ? ((t & 1) | ((t >> 2) & 2) | ((t >> 4) & 4))
         : d == 1 ? (((t >> 1) & 1) | ((t >> 3) & 2) | ((t >> 5) & 4))
                  : (((t >> 2) & 1) | ((t >> 4) & 2));
  }
  T8B_HD static int index(int x, int y, int z) {
    return (x & 1) | (y & 1) << 1 | (z & 1) << 2 | (x & 2) << 2 | (y & 2) << 3 | (z & 2) << 4 | (x & 4) << 4 | (y & 4) << 5;
  }
  T8B_HD static bool at_lower(int t, int d) { return (t & mask(d)) == 0; }
  T8B_HD static bool at_upper(int t, int d) { return (t & mask(d)) == mask(d); }
  // own-slot neighbours (only meaningful away from the boundary)
  T8B_HD static int lower_own(int t, int d) { return (((t & mask(d)) - 1) & mask(d)) | (t & ~mask(d)); }
  T8B_HD static int upper_own(int t, int d) { return (((t | ~mask(d)) + 1) & mask(d)) | (t & ~mask(d)); }
  // index of boundary element t inside its boundary plane
  T8B_HD static int compact(int t, int d) {
    return d == 0 ? (((t >> 1) & 3) | ((t >> 4) & 3) << 2 | ((t >> 7) & 1) << 4)
         : d == 1 ? ((t & 1) | ((t >> 2) & 3) << 1 | ((t >> 5) & 3) << 3)
                  : ((t & 3) | ((t >> 3) & 3) << 2 | ((t >> 6) & 3) << 4);
  }
  // element on the upper boundary plane with compact index idx
  T8B_HD static int upper_elem(int idx, int d) {
    return d == 0 ? (0x49 | (idx & 3) << 1 | ((idx >> 2) & 3) << 4 | ((idx >> 4) & 1) << 7)
         : d == 1 ? (0x92 | (idx & 1) | ((idx >> 1) & 3) << 2 | ((idx >> 3) & 3) << 5)
                  : (0x24 | (idx & 3) | ((idx >> 2) & 3) << 3 | ((idx >> 4) & 3) << 6);
  }
  T8B_HD static int halo_slot(int d, int side, int idx) {
    return d == 0 ? 256 + 16 * (idx >> 2) + 8 + (side ? 0 : 1) + 2 * (idx & 3)
         : d == 1 ? 256 + 16 * (8 + (idx >> 3)) + ((idx & 1) | (side ? 0 : 2) | ((idx >> 1) & 3) << 2)
                  : 256 + 16 * (12 + (idx >> 3)) + ((idx & 3) | (side ? 0 : 4) | ((idx >> 2) & 1) << 3);
  }
  T8B_HD static int thread_slot(int h) { return h < 64 ? 264 + 16 * (h >> 3) + (h & 7) : 320 + h; }
};

struct SubgridBox {
  static constexpr int NSLOT = 256 + 384;
  T8B_HD static constexpr unsigned mask(int d) { return d == 0 ? 0x43u : d == 1 ? 0x8Cu : 0x30u; }
  T8B_HD static constexpr int      wrap(int d) { return d == 0 ? 61 : d == 1 ? 116 : 0; }   // into the sibling element: t - 64 + 3, t - 128 + 12
  T8B_HD static constexpr int      step(int d) { return d == 0 ? 1 : d == 1 ? 4 : 16; }
  T8B_HD static constexpr unsigned low(int d) { return d == 0 ? 0x03u : d == 1 ? 0x0Cu : 0x30u; }   // bits of the in-element coordinate
  T8B_HD static int coord(int t, int d) {
    return d == 0 ? (4 * ((t >> 6) & 1) + (t & 3)) : d == 1 ? (4 * (t >> 7) + ((t >> 2) & 3)) : ((t >> 4) & 3);
  }
  T8B_HD static int index(int x, int y, int z) { return ((x >> 2) + 2 * (y >> 2)) * 64 + (x & 3) + 4 * (y & 3) + 16 * z; }
  T8B_HD static bool at_lower(int t, int d) { return (t & mask(d)) == 0; }
  T8B_HD static bool at_upper(int t, int d) { return (t & mask(d)) == mask(d); }
  T8B_HD static int lower_own(int t, int d) { return (t & low(d)) ? t - step(d) : t - wrap(d); }
  T8B_HD static int upper_own(int t, int d) { return (t & low(d)) != low(d) ? t + step(d) : t + wrap(d); }
  T8B_HD static int compact(int t, int d) {
    return d == 0 ? (((t >> 2) & 3) | ((t >> 4) & 3) << 2 | ((t >> 7) & 1) << 4)
         : d == 1 ? ((t & 3) | ((t >> 4) & 3) << 2 | ((t >> 6) & 1) << 4)
                  : ((t & 15) | ((t >> 6) & 3) << 4);
  }
  T8B_HD static int upper_elem(int idx, int d) {
    return d == 0 ? (64 + 3 + 4 * (idx & 3) + 16 * ((idx >> 2) & 3) + 128 * (idx >> 4))
         : d == 1 ? (128 + 12 + (idx & 3) + 16 * ((idx >> 2) & 3) + 64 * (idx >> 4))
                  : (48 + (idx & 15) + 64 * (idx >> 4));
  }
  T8B_HD static int halo_slot(int d, int side, int idx) {
    return d == 0 ? 256 + 16 * (idx >> 2) + (side ? 0 : 3) + 4 * (idx & 3)
         : d == 1 ? 256 + 16 * (8 + (idx >> 2)) + (side ? 0 : 12) + (idx & 3)
                  : 256 + 16 * ((side ? 20 : 16) + (idx >> 4)) + (idx & 15);
  }
  T8B_HD static int thread_slot(int h) {
    return h < 64 ? 256 + 16 * (h >> 3) + 4 * ((h & 7) >> 1) + ((h & 1) ? 3 : 0)
         : h < 128 ? 384 + 16 * ((h - 64) >> 3) + ((h & 7) < 4 ? (h & 7) : 8 + (h & 7))
                   : 384 + h;
  }
};

// shared by both layouts
struct BoxCommon {
  static constexpr int NFLUX = 768 + 128;
  T8B_HD static int upper_flux(int d, int idx) { return 768 + (d < 2 ? 32 * d : 64) + idx; }
};

}  // namespace t8b200
