// Entropy-stable KEPES flux with matrix dissipation for the compressible Euler equations (gamma = 1.4),
// written for sm_100a.  Replaces (not a translation of):
//   examples/compressible_euler/kernels.cu:24-133,174-290   (ln_mean, kepes_compute_flux, diffusion matrix, rotation)
//   examples/subgrid/kernels.inl:21-261
//
// Design differences from the reference (all algebraically exact, so results agree to rounding):
//  * rotation-free: the reference builds an orthonormal frame (n,t1,t2) per face (1 sqrt + 3 divides), rotates both
//    states, evaluates the flux in that frame and rotates back.  Every term of F* and of R D R^T (wR - wL) is either
//    rotation invariant or a multiple of n / of a vector already known in xyz, so we evaluate directly in xyz with
//    dot products against n.
//  * per-cell primitives (rho, v, p, beta = rho/2p) are computed once per cell per stage and staged in shared memory
//    instead of being recomputed (with 4 divides) by each of the cell's faces.
//  * no log() for the entropy-variable jump: s = log p - kappa log rho is only needed as sR - sL, and
//    log(aR/aL) = (aR - aL) / ln_mean(aL, aR) is a by-product of the two logarithmic means the flux needs anyway
//    (series branch: 2 f (1 + u/3 + u^2/5 + u^3/7), f = (aR-aL)/(aR+aL), u = f^2).  The reference calls log 4x per face.
//  * the sparse eigenvector matrix R (11 of 25 entries are 0/1) is expanded by hand.
#pragma once
#include <cuda_runtime.h>

#define T8B_HD __host__ __device__ __forceinline__

namespace t8b200 {

template <typename T>
struct Prim {
  T rho, vx, vy, vz, p, beta;
};

template <typename T>
T8B_HD T fabs_(T x) { return x < T(0) ? -x : x; }
template <typename T>
T8B_HD T fmax_(T a, T b) { return a > b ? a : b; }

// conserved (rho, m, E) -> primitives.  kernels.cu:54-71.
template <typename T>
T8B_HD Prim<T> to_prim(T rho, T mx, T my, T mz, T e) {
  Prim<T> q;
  const T km1 = T(1.4) - T(1);
  T       sr  = T(1) / rho;
  q.rho       = rho;
  q.vx        = sr * mx;
  q.vy        = sr * my;
  q.vz        = sr * mz;
  T ke        = T(0.5) * (q.vx * q.vx + q.vy * q.vy + q.vz * q.vz);
  q.p         = km1 * (e - rho * ke);
  q.beta      = T(0.5) * rho / q.p;
  return q;
}

// Logarithmic mean and log(aR/aL) in one go.  kernels.cu:24-36 (Ismail-Roe).
template <typename T>
T8B_HD void ln_mean_and_log(T aL, T aR, T& mean, T& lograt) {
  T s = aL + aR;
  T d = aR - aL;
  T f = d / s;
  T u = f * f;
  if (u < T(1.0e-4)) {
    // 105 + 35u + 21u^2 + 15u^3 = 105 (1 + u/3 + u^2/5 + u^3/7)
    T P    = T(105.0) + u * (T(35.0) + u * (T(21.0) + u * T(15.0)));
    mean   = s * T(52.5) / P;
    lograt = f * P * T(2.0 / 105.0);
  } else {
    lograt = log(aR / aL);
    mean   = d / lograt;
  }
}

// Numerical flux through a face with unit normal n (pointing L -> R), in xyz, NOT scaled by the area.
// Returns the wave-speed estimate |uHat| + aHat (kernels.cu:222).
template <typename T>
T8B_HD T kepes_flux(const Prim<T>& L, const Prim<T>& R, T nx, T ny, T nz, T F[5]) {
  const T kappa = T(1.4);
  const T km1   = kappa - T(1);
  const T half  = T(0.5);

  T rhoHat, dlogrho, betaHat, dlogbeta;
  ln_mean_and_log(L.rho, R.rho, rhoHat, dlogrho);
  ln_mean_and_log(L.beta, R.beta, betaHat, dlogbeta);

  T rhoMean  = half * (L.rho + R.rho);
  T betaMean = half * (L.beta + R.beta);
  T ax = half * (L.vx + R.vx), ay = half * (L.vy + R.vy), az = half * (L.vz + R.vz);  // averaged velocity
  T pMean = half * (L.p + R.p);

  T qL = half * (L.vx * L.vx + L.vy * L.vy + L.vz * L.vz);
  T qR = half * (R.vx * R.vx + R.vy * R.vy + R.vz * R.vz);

  T uHat  = ax * nx + ay * ny + az * nz;
  T vv    = ax * ax + ay * ay + az * az;
  T aHat  = sqrt(kappa * pMean / rhoHat);
  T ib    = T(1) / betaHat;
  T HHat  = (kappa / (T(2) * km1)) * ib + half * (L.vx * R.vx + L.vy * R.vy + L.vz * R.vz);
  T p1Hat = half * rhoMean / betaMean;

  // entropy-conservative part, kernels.cu:86-92
  T F0  = rhoHat * uHat;
  T Fsx = F0 * ax + p1Hat * nx;
  T Fsy = F0 * ay + p1Hat * ny;
  T Fsz = F0 * az + p1Hat * nz;
  T Fs4 = F0 * (half * (ib / km1 - (qL + qR)) + vv) + uHat * p1Hat;

  // jump of the entropy variables, kernels.cu:227-266
  T bL2 = L.beta + L.beta, bR2 = R.beta + R.beta;  // rho/p
  T J0  = dlogrho + dlogbeta / km1 - (bR2 * qR - bL2 * qL);
  T Jx = bR2 * R.vx - bL2 * L.vx, Jy = bR2 * R.vy - bL2 * L.vy, Jz = bR2 * R.vz - bL2 * L.vz;
  T J4 = bL2 - bR2;

  T vJ = ax * Jx + ay * Jy + az * Jz;
  T Jn = nx * Jx + ny * Jy + nz * Jz;
  T g  = Jn + uHat * J4;

  // R^T J, scaled by D (kernels.cu:114-132, 267-270)
  T c  = J0 + vJ + HHat * J4;
  T e  = aHat * g;
  T a1 = J0 + vJ + half * vv * J4;
  T rk = rhoHat / kappa;
  T d0 = half * fabs_(uHat - aHat) * rk * (c - e);
  T d4 = half * fabs_(uHat + aHat) * rk * (c + e);
  T au = fabs_(uHat);
  T d1 = au * km1 * rk * a1;
  T D2 = au * p1Hat;

  // tangential part of (Jm + vbar J4)
  T tx = Jx + ax * J4 - nx * g, ty = Jy + ay * J4 - ny * g, tz = Jz + az * J4 - nz * g;

  // R (D R^T J), kernels.cu:272-275
  T sum = d0 + d1 + d4;
  T dif = aHat * (d4 - d0);
  T ds4 = HHat * (d0 + d4) + uHat * dif + half * vv * d1 + D2 * (vJ + vv * J4 - uHat * g);

  F[0] = F0 - half * sum;
  F[1] = Fsx - half * (sum * ax + dif * nx + D2 * tx);
  F[2] = Fsy - half * (sum * ay + dif * ny + D2 * ty);
  F[3] = Fsz - half * (sum * az + dif * nz + D2 * tz);
  F[4] = Fs4 - half * ds4;
  return au + aHat;
}

// wall boundary: right state = left state with the normal velocity mirrored (kernels.cu:371-375)
template <typename T>
T8B_HD Prim<T> mirror(const Prim<T>& L, T nx, T ny, T nz) {
  Prim<T> R = L;
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

}  // namespace t8b200
