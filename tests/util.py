"""Shared helpers for the parity tests (test infrastructure; may import the oracle)."""
import numpy as np

import oracle

TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}  # north_star tolerances, per step


def rel_linf(a, b):
    """Relative L-infinity difference per conserved variable, worst over variables.  A variable whose reference
    magnitude is negligible against the state scale (e.g. rho_v2 == 0 in the 3-D KH setup) is measured against
    the momentum/state scale instead of its own ~0 maximum."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if not (np.isfinite(a).all() and np.isfinite(b).all()):
        return float("inf")          # a NaN must never compare as "equal"
    scale_all = np.abs(b).max()
    worst = 0.0
    for k in range(a.shape[0]):
        sk = np.abs(b[k]).max()
        if sk < 1e-3 * scale_all:
            sk = scale_all
        worst = max(worst, np.abs(a[k] - b[k]).max() / sk)
    return worst


def perturbed_kh(forest, dtype, seed=0, amp=0.05):
    """Cartesian KH state at the centroids plus a smooth-ish seeded perturbation (so that no flux term vanishes)."""
    lv, cent, vol, _ = forest.elements()
    u = oracle.init_kh_points(forest.dim, cent.astype(dtype), dtype).astype(np.float64)
    rng = np.random.default_rng(seed)
    n = u.shape[1]
    rho = u[0] * (1 + amp * rng.uniform(-1, 1, n))
    v = u[1:4] / u[0] + amp * rng.uniform(-1, 1, (3, n))
    if forest.dim == 2:
        v[2] = 0.0
    p = 2.5 * (1 + amp * rng.uniform(-1, 1, n))
    out = np.empty_like(u)
    out[0] = rho
    out[1:4] = rho * v
    out[4] = p / 0.4 + 0.5 * rho * (v * v).sum(0)
    return np.ascontiguousarray(out.astype(dtype)), vol.astype(dtype)


def global_reference_steps(forest, u0, vol, dt, nsteps, dtype):
    """Oracle: nsteps of iterate() on the single-rank connectivity; returns list of states after each step."""
    conn = forest.connectivity(1, 0, dtype=dtype)
    out, u = [], u0
    for _ in range(nsteps):
        u, _, _ = oracle.iterate(conn, vol, u, dt)
        out.append(u)
    return out


# ------------------------------------------------------------------------------------------------ hybrid meshes
from t8gpu_b200.meshes import hybrid_mesh, partition_flat_mesh, smooth_state, tile_periodic_mesh  # noqa: E402,F401
