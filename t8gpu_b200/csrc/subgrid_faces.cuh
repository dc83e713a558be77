// Face source of the cell-level plans of Subgrid<4,4,4> / Subgrid<4,4>, callable from the host builder (tile_plan.cuh)
// and from the device builder (device_plan.cu, plan_block.cuh).
#pragma once
#include <cmath>
#include <cstdint>

#include "common.cuh"
#include "euler_flux.cuh"

// cell index inside an element from (axis, x along the axis, tangential a, b)
T8B_HD int cell_ax(int dim, int ax, int x, int a, int b) {
  if (dim == 2) return ax == 0 ? x + 4 * a : a + 4 * x;
  return ax == 0 ? x + 4 * a + 16 * b : (ax == 1 ? a + 4 * x + 16 * b : a + 4 * b + 16 * x);
}

// DIM is a template parameter: the plan builder asks for the endpoints of every cell face several times, and with
// compile-time cell / face counts the decoding of a face id needs no integer division by a run-time value
template <typename T, int DIM>
struct SubgridFaces {
  static constexpr int dim = DIM;
  int64_t        n_local;  // elements
  int32_t        nf, nb, nx;
  const int32_t* nbr;
  const T *      normals, *areas;   // normals: dim components per face
  const int32_t *ld, *off;          // off: dim components per face
  const T*       vol;
  const int32_t *ranks, *indices, *xnbr;
  const T *      xnormals, *xareas;
  const int32_t *xld, *xoff;
  const T*       inner_area;   // per element: area of the faces between its cells (subgrid_inner_area)
  T8B_HD static constexpr int S() { return DIM == 3 ? 64 : 16; }
  T8B_HD static constexpr int TPF() { return DIM == 3 ? 16 : 4; }
  T8B_HD static constexpr int IPE() { return DIM * 3 * TPF(); }   // inner faces per element: dim axes x 3 planes x TPF
  T8B_HD int64_t n_inner() const { return n_local * IPE(); }
  T8B_HD int64_t num_faces() const { return n_inner() + ((int64_t)nf + nb + nx) * TPF(); }

  // element face F (0..nf+nb+nx), sub-face s -> left / right cell inside their elements
  T8B_HD void sub_cells(const T* n, const int32_t* o, int dstride, int s, int& lc, int& rc) const {
    const int i = s & 3, j = s >> 2;
    int al[3] = {0, 0, 0}, si[3] = {0, 0, 0}, sj[3] = {0, 0, 0};
    if (n[0] == T(1)) { al[0] = 3; si[1] = 1; sj[2] = 1; }
    if (n[0] == T(-1)) { si[1] = 1; sj[2] = 1; }
    if (n[1] == T(1)) { al[1] = 3; si[0] = 1; sj[2] = 1; }
    if (n[1] == T(-1)) { si[0] = 1; sj[2] = 1; }
    if (dim == 3) {
      if (n[2] == T(1)) { al[2] = 3; si[0] = 1; sj[1] = 1; }
      if (n[2] == T(-1)) { si[0] = 1; sj[1] = 1; }
    }
    int l[3], r[3];
    for (int d = 0; d < 3; d++) {
      l[d] = al[d] + i * si[d] + j * sj[d];
      r[d] = (o && d < dim ? o[d] : 0) + dstride * (i * si[d] + j * sj[d]) / 2;
    }
    lc = l[0] + 4 * l[1] + 16 * l[2];
    rc = r[0] + 4 * r[1] + 16 * r[2];
  }
  T8B_HD void outer(int64_t g, int64_t& F, int& s, const int32_t*& pn, const T*& n, const T*& ar, const int32_t*& l_d,
             const int32_t*& o, bool& wall) const {
    F = g / TPF();
    s = (int)(g % TPF());
    wall = false;
    if (F < nf) { pn = nbr + 2 * F; n = normals + dim * F; ar = areas + F; l_d = ld + F; o = off + dim * F; }
    else if (F < (int64_t)nf + nb) {
      wall = true;
      pn = nbr + 2 * (int64_t)nf + (F - nf); n = normals + dim * F; ar = areas + F; l_d = nullptr; o = nullptr;
    } else {
      const int64_t x = F - nf - nb;
      pn = xnbr + 2 * x; n = xnormals + dim * x; ar = xareas + x; l_d = xld + x; o = xoff + dim * x;
    }
  }
  T8B_HD void endpoints(int64_t f, int32_t& l, int32_t& r) const {
    if (f < n_inner()) {
      const int64_t e = f / IPE();
      const int     q = (int)(f % IPE()), ax = q / (3 * TPF()), t = q % (3 * TPF()), p = t / TPF(), s = t % TPF();
      l = (int32_t)(e * S() + cell_ax(dim, ax, p, s & 3, s >> 2));
      r = (int32_t)(e * S() + cell_ax(dim, ax, p + 1, s & 3, s >> 2));
      return;
    }
    int64_t        F;
    int            s, lc, rc;
    const int32_t *pn, *l_d, *o;
    const T *      n, *ar;
    bool           wall;
    outer(f - n_inner(), F, s, pn, n, ar, l_d, o, wall);
    sub_cells(n, o, (l_d && *l_d != 0) ? 1 : 2, s, lc, rc);
    l = (int32_t)((int64_t)pn[0] * S() + lc);
    r = wall ? -1 : (int32_t)((int64_t)pn[1] * S() + rc);
  }
  T8B_HD void geometry(int64_t f, T nrm[3], T& a) const {
    nrm[0] = nrm[1] = nrm[2] = T(0);
    if (f < n_inner()) {
      const int64_t e  = f / IPE();
      const int     ax = (int)(f % IPE()) / (3 * TPF());
      nrm[ax] = T(1);
      a = inner_area[e];
      return;
    }
    int64_t        F;
    int            s;
    const int32_t *pn, *l_d, *o;
    const T *      n, *ar;
    bool           wall;
    outer(f - n_inner(), F, s, pn, n, ar, l_d, o, wall);
    for (int d = 0; d < dim; d++) nrm[d] = n[d];
    a = *ar / T(TPF());
  }
  T8B_HD void owner(int32_t id, int32_t& rk, int32_t& ix) const {
    const int32_t e = id / S(), c = id % S();
    rk = ranks[e];
    ix = indices[e] * S() + c;
  }
};


// area of the faces between the cells of one element: (cbrt(vol)/4)^2 (kernels.inl:352-354) resp. sqrt(vol)/4 (2-D,
// :542-544).  cbrt is off by an ulp on exact cubes (dyadic Cartesian volumes) in libm and in CUDA's libdevice alike: the
// exact root is taken when there is one, so that the faces inside an element and between elements get the same area
// entry -- and the host and the device builder the same bits
template <typename T, int DIM>
T8B_HD T subgrid_inner_area(T vol) {
  if (DIM == 3) {
    T       c  = cbrt(vol);
    const T lo = nextafter(c, T(0)), hi = nextafter(c, T(2) * c);
    if (lo * lo * lo == vol) c = lo;
    if (hi * hi * hi == vol) c = hi;
    const T edge = c / T(4);
    return edge * edge;
  }
  return sqrt(vol) / T(4);
}
