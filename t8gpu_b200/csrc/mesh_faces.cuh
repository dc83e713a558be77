// Face source of the unstructured plans, callable from the host builder (tile_plan.cuh) and from the device builder
// (device_plan.cu, plan_block.cuh).
#pragma once
#include <cstdint>

#include "common.cuh"
#include "euler_flux.cuh"

// Face source over the arrays behind MeshConnectivityAccessor<float_type,3> (t8gpu/mesh/mesh_manager.h:159-166):
// global face ids [0,nf) interior, [nf,nf+nb) boundary, then the extra partition-boundary faces.
template <typename T>
struct MeshFaces {
  int32_t        nf, nb, nx;
  const int32_t* nbr;
  const T *      normals, *areas;
  const int32_t *ranks, *indices, *xnbr;
  const T *      xnormals, *xareas;
  T8B_HD int64_t num_faces() const { return (int64_t)nf + nb + nx; }
  T8B_HD void endpoints(int64_t f, int32_t& l, int32_t& r) const {
    if (f < nf) { l = nbr[2 * f]; r = nbr[2 * f + 1]; }
    else if (f < (int64_t)nf + nb) { l = nbr[2 * (int64_t)nf + (f - nf)]; r = -1; }
    else { int64_t g = f - nf - nb; l = xnbr[2 * g]; r = xnbr[2 * g + 1]; }
  }
  T8B_HD void geometry(int64_t f, T nrm[3], T& a) const {
    const T* n;
    if (f < (int64_t)nf + nb) { n = normals + 3 * f; a = areas[f]; }
    else { int64_t g = f - nf - nb; n = xnormals + 3 * g; a = xareas[g]; }
    nrm[0] = n[0]; nrm[1] = n[1]; nrm[2] = n[2];
  }
  T8B_HD void owner(int32_t id, int32_t& rk, int32_t& ix) const { rk = ranks[id]; ix = indices[id]; }
};

