// Entropy-stable KEPES flux with matrix dissipation for the compressible Euler equations (gamma = 1.4),
// written for sm_100a.  Replaces (not a translation of):
//   examples/compressible_euler/kernels.cu:24-133,174-290   (ln_mean, kepes_compute_flux, diffusion matrix, rotation)
//   examples/subgrid/kernels.inl:21-261
//
// The reference spends ~680 FP64-pipe instructions per face, which on B200 (64 FP64 lanes / SM /
// clock) is ~3x above the HBM floor.  This formulation needs ~150, all algebraically exact rewrites (results agree
// with the reference's evaluation order to rounding, far inside the 1e-12 / 1e-5 tolerances):
//  * rotation-free: the reference builds an orthonormal frame (n,t1,t2) per face (1 sqrt + 3 divides), rotates both
//    states, evaluates the flux in that frame and rotates back.  Every term of F* and of R D R^T (wR - wL) is either
//    rotation invariant or a multiple of n / of a vector already known in xyz, so we evaluate directly in xyz with
//    dot products against n.
//  * per-cell quantities (rho, v, p, B = rho/p = 2 beta, w = B |v|^2 / 2) are computed once per cell per stage
//    (one reciprocal) and staged in shared memory instead of being recomputed, with 4 divides, by each face.
//  * no log() for the entropy-variable jump: s = log p - kappa log rho is only needed as sR - sL, and
//    log(aR/aL) = (aR - aL) / ln_mean(aL, aR) is a by-product of the two logarithmic means the flux needs anyway.
//    The reference calls log 4x per face (kernels.cu:236-237).
//  * the Ismail-Roe series branch of ln_mean, (aL+aR) * 52.5 / (105 + 35u + 21u^2 + 15u^3) with u = f^2 < 1e-4,
//    f = (aR-aL)/(aR+aL), is evaluated without its divide: with x = u/3 + u^2/5 + u^3/7 <= 3.4e-5,
//    mean = (s/2)(1 - x + x^2 - x^3), 1/mean = (2/s)(1 + x), log(aR/aL) = 2 f (1 + x); truncation x^4 < 2e-18.
//  * the two remaining reciprocals per face, 1/(rhoL+rhoR) and 1/(BL+BR), come from ONE reciprocal of their product;
//    1/betaMean is 4/(BL+BR) (same quantity), so p1Hat = (rhoL+rhoR)/(BL+BR).
//  * Fs4's  1/2 (1/((k-1) betaHat) - |vL|^2/2 - |vR|^2/2) + |vbar|^2  collapses to 1/(2 (k-1) betaHat) + vL.vR/2.
//  * reciprocal / rsqrt seeds from MUFU (rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64) + 2 Newton steps: no slow-path
//    branches; relative error <= ~2 ulp.
//  * the sparse eigenvector matrix R (11 of 25 entries are 0/1) is expanded by hand.
#pragma once
#include <cuda_runtime.h>

#include <cmath>

#define T8B_HD __host__ __device__ __forceinline__

namespace t8b200 {

// Per-cell quantities staged in shared memory (7 values).
template <typename T>
struct Cell {
  T rho, vx, vy, vz, p, B, w;  // B = rho / p ( = 2 beta ),  w = B * |v|^2 / 2
};
constexpr int NCELLQ = 7;

template <typename T>
T8B_HD T fabs_(T x) { return x < T(0) ? -x : x; }
template <typename T>
T8B_HD T fmax_(T a, T b) { return a > b ? a : b; }

// ---- fast reciprocal / rsqrt ------------------------------------------------------------------------------
T8B_HD double fast_rcp(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r        = fma(r, e, r);
  e        = fma(-x, r, 1.0);
  r        = fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}
T8B_HD float fast_rcp(float x) {
#ifdef __CUDA_ARCH__
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}
T8B_HD double fast_sqrt(double x) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double xh = 0.5 * x;
  double t  = fma(-xh * y, y, 0.5);
  y         = fma(y, t, y);
  t         = fma(-xh * y, y, 0.5);
  y         = fma(y, t, y);
  double s  = x * y;
  return fma(fma(-s, s, x), 0.5 * y, s);
#else
  return std::sqrt(x);
#endif
}
T8B_HD float fast_sqrt(float x) {
#ifdef __CUDA_ARCH__
  return __fsqrt_rn(x);
#else
  return std::sqrt(x);
#endif
}

// conserved (rho, m, E) -> per-cell quantities.  kernels.cu:54-71, 230-240.
template <typename T>
T8B_HD Cell<T> to_cell(T rho, T mx, T my, T mz, T e) {
  Cell<T> q;
  const T km1 = T(1.4) - T(1);
  q.rho = rho;
  // p = (k-1)(e - |m|^2 / (2 rho)) ; one reciprocal serves 1/rho and 1/p
  T m2  = mx * mx + my * my + mz * mz;
  T pr  = km1 * (e * rho - T(0.5) * m2);  // = p * rho
  T r   = fast_rcp(pr * rho);             // 1 / (p rho^2)
  T sr  = r * pr;                         // 1 / rho = (p rho) / (p rho^2)
  T sp  = r * rho * rho;                  // 1 / p   = rho^2 / (p rho^2)
  q.vx  = sr * mx;
  q.vy  = sr * my;
  q.vz  = sr * mz;
  q.p   = pr * sr;
  q.B   = rho * sp;
  q.w   = T(0.5) * q.B * (q.vx * q.vx + q.vy * q.vy + q.vz * q.vz);
  return q;
}

// mean = ln_mean(aL,aR) (kernels.cu:24-36), imean = 1/mean, lograt = log(aR/aL), given s = aL+aR, d = aR-aL and
// is = 1/s.
template <typename T>
T8B_HD void ln_mean3(T aL, T aR, T s, T d, T is, T& mean, T& imean, T& lograt) {
  T f = d * is;
  T u = f * f;
  if (u < T(1.0e-4)) {
    T x    = u * (T(1.0 / 3.0) + u * (T(1.0 / 5.0) + u * T(1.0 / 7.0)));
    T y    = T(1) + x;
    mean   = T(0.5) * s * (T(1) - x * (T(1) - x * (T(1) - x)));
    imean  = (is + is) * y;
    lograt = (f + f) * y;
  } else {
    lograt = log(aR / aL);
    mean   = d / lograt;
    imean  = lograt / d;
  }
}

// Numerical flux through a face with unit normal n (pointing L -> R), in xyz, NOT scaled by the area.
// Returns the wave-speed estimate |uHat| + aHat (kernels.cu:222).
template <typename T>
T8B_HD T kepes_flux(const Cell<T>& L, const Cell<T>& R, T nx, T ny, T nz, T F[5]) {
  const T kappa = T(1.4);
  const T km1   = kappa - T(1);
  const T half  = T(0.5);

  T sr = L.rho + R.rho, dr = R.rho - L.rho;
  T sB = L.B + R.B, dB = R.B - L.B;
  T rr  = fast_rcp(sr * sB);
  T isr = rr * sB, isB = rr * sr;

  T rhoHat, irhoHat, dlogrho, BHat, iBHat, dlogB;
  ln_mean3(L.rho, R.rho, sr, dr, isr, rhoHat, irhoHat, dlogrho);
  ln_mean3(L.B, R.B, sB, dB, isB, BHat, iBHat, dlogB);
  (void)BHat;

  T ax = half * (L.vx + R.vx), ay = half * (L.vy + R.vy), az = half * (L.vz + R.vz);  // averaged velocity
  T uHat = ax * nx + ay * ny + az * nz;
  T vv   = ax * ax + ay * ay + az * az;
  T dLR  = half * (L.vx * R.vx + L.vy * R.vy + L.vz * R.vz);
  T aHat = fast_sqrt((kappa * half) * (L.p + R.p) * irhoHat);
  // 1/betaHat = 2/BHat ;  kappa/(2 (k-1) betaHat) = kappa/(k-1) * iBHat
  T ibk   = iBHat / km1;            // 1 / (2 (k-1) betaHat)
  T HHat  = kappa * ibk + dLR;      // kernels.cu:82
  T p1Hat = sr * isB;               // (rhoMean/2) / betaMean, kernels.cu:83

  // entropy-conservative part, kernels.cu:86-92
  T F0  = rhoHat * uHat;
  T Fs4 = F0 * (ibk + dLR) + uHat * p1Hat;

  // jump of the entropy variables, kernels.cu:227-266
  T J0 = dlogrho + dlogB / km1 - (R.w - L.w);
  T Jx = R.B * R.vx - L.B * L.vx, Jy = R.B * R.vy - L.B * L.vy, Jz = R.B * R.vz - L.B * L.vz;
  T J4 = -dB;

  T vJ = ax * Jx + ay * Jy + az * Jz;
  T Jn = nx * Jx + ny * Jy + nz * Jz;
  T g  = Jn + uHat * J4;

  // R^T J, scaled by D (kernels.cu:114-132, 267-270)
  T b  = J0 + vJ;
  T c  = b + HHat * J4;
  T e  = aHat * g;
  T a1 = b + half * vv * J4;
  T rk = rhoHat * (half / kappa);
  T d0 = fabs_(uHat - aHat) * rk * (c - e);
  T d4 = fabs_(uHat + aHat) * rk * (c + e);
  T au = fabs_(uHat);
  T d1 = au * (T(2) * km1) * rk * a1;
  T D2 = au * p1Hat;

  // R (D R^T J), kernels.cu:272-275.  t = tangential part of (Jm + vbar J4).
  T sum = d0 + d1 + d4;
  T dif = aHat * (d4 - d0);
  T ds4 = HHat * (d0 + d4) + uHat * dif + half * vv * d1 + D2 * (vJ + vv * J4 - uHat * g);
  // momentum: Fs_m - 1/2 (sum vbar + dif n + D2 (Jm + vbar J4 - n g))  =  (F0 - sum/2 - D2 J4/2) vbar
  //                                                                        + (p1Hat - dif/2 + D2 g/2) n - D2/2 Jm
  T hD2 = half * D2;
  T ca  = F0 - half * sum - hD2 * J4;
  T cn  = p1Hat - half * dif + hD2 * g;

  F[0] = F0 - half * sum;
  F[1] = ca * ax + cn * nx - hD2 * Jx;
  F[2] = ca * ay + cn * ny - hD2 * Jy;
  F[3] = ca * az + cn * nz - hD2 * Jz;
  F[4] = Fs4 - half * ds4;
  return au + aHat;
}

// wall boundary: right state = left state with the normal velocity mirrored (kernels.cu:371-375)
template <typename T>
T8B_HD Cell<T> mirror(const Cell<T>& L, T nx, T ny, T nz) {
  Cell<T> R  = L;
  T       vn = L.vx * nx + L.vy * ny + L.vz * nz;
  R.vx -= T(2) * vn * nx;
  R.vy -= T(2) * vn * ny;
  R.vz -= T(2) * vn * nz;
  return R;
}

// SSP-RK3 stage combination, ssp_runge_kutta.inl:3-26,43,66-68,91-93 (truncated literals and evaluation order kept).
template <typename T, int STAGE>
T8B_HD T rk_combine(T prev, T in, T flux, T dt, T vol) {
  if (STAGE == 1) return prev + dt / vol * flux;
  if (STAGE == 2) return T(0.75) * prev + T(0.25) * in + T(0.25) * dt / vol * flux;
  return T(0.33333333333333) * prev + T(0.66666666666666) * in + T(0.66666666666666) * dt / vol * flux;
}
// runtime-stage variant with the scale factor c*dt/vol hoisted by the caller (fused kernels)
template <typename T>
T8B_HD T rk_scale(int stage, T dt, T vol) {
  if (stage == 1) return dt / vol;
  if (stage == 2) return T(0.25) * dt / vol;
  return T(0.66666666666666) * dt / vol;
}
template <typename T>
T8B_HD T rk_apply(int stage, T prev, T in, T flux, T scale) {
  if (stage == 1) return prev + scale * flux;
  if (stage == 2) return T(0.75) * prev + T(0.25) * in + scale * flux;
  return T(0.33333333333333) * prev + T(0.66666666666666) * in + scale * flux;
}

#ifdef __CUDACC__
// block-wide max of a non-negative value, then one atomicMax on its bit pattern (order-preserving for x >= 0).
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T w = __shfl_xor_sync(0xffffffffu, v, o);
    v   = v > w ? v : w;
  }
  return v;
}
#endif

}  // namespace t8b200
