// TEST-ONLY: evaluates the product's flux algebra (t8gpu_b200/csrc/euler_flux.cuh) on the host so the "not gpu"
// suite can compare it against the oracle without a GPU.  Not part of the product library.
#include "../../t8gpu_b200/csrc/euler_flux.cuh"
using namespace t8b200;
template <typename T>
static T run(const T* uL, const T* uR, const T* n, int reflect, T* F) {
  Cell<T> L = to_cell(uL[0], uL[1], uL[2], uL[3], uL[4]);
  Cell<T> R = reflect ? mirror(L, n[0], n[1], n[2]) : to_cell(uR[0], uR[1], uR[2], uR[3], uR[4]);
  // axis-aligned normals +e_k also exercise the specialised variant the fused kernels use
  for (int k = 0; k < 3; k++)
    if (n[k] == T(1) && n[(k + 1) % 3] == T(0) && n[(k + 2) % 3] == T(0)) {
      if (k == 0) return kepes_flux_n<T, 0>(L, R, T(0), T(0), T(0), F);
      if (k == 1) return kepes_flux_n<T, 1>(L, R, T(0), T(0), T(0), F);
      return kepes_flux_n<T, 2>(L, R, T(0), T(0), T(0), F);
    }
  return kepes_flux(L, R, n[0], n[1], n[2], F);
}
extern "C" double hostcheck_flux_f64(const double* uL, const double* uR, const double* n, int reflect, double* F) {
  return run<double>(uL, uR, n, reflect, F);
}
extern "C" float hostcheck_flux_f32(const float* uL, const float* uR, const float* n, int reflect, float* F) {
  return run<float>(uL, uR, n, reflect, F);
}
