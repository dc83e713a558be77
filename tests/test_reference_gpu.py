"""GPU parity against the reference's OWN CUDA implementation (oracle/_ref: examples/compressible_euler and
examples/subgrid compiled unmodified from the reference sources for sm_100a, running over the t8mini forest facade).
Pins (i) the oracle's arithmetic, (ii) the oracle's connectivity restatement against the reference's own
compute_connectivity_information / adapt code, (iii) the product against the reference itself."""
import numpy as np
import pytest
import torch

import oracle
from oracle import ref_cuda
from util import TOL, perturbed_kh, rel_linf

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built (needs /root/reference)")]

DT = {np.float64: torch.float64, np.float32: torch.float32}
CONN_KEYS = ("ranks", "indices", "face_neighbors", "face_normals", "face_areas")


def assert_conn_equal(ref, orc, subgrid=False):
    for k in ("n_local", "n_ghost", "n_faces", "n_bfaces"):
        assert ref[k] == orc[k], k
    keys = CONN_KEYS + (("level_diff", "offsets") if subgrid else ())
    for k in keys:
        assert ref[k].dtype == orc[k].dtype, k
        assert np.array_equal(ref[k], orc[k]), k


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,level,periodic", [(2, 6, True), (3, 3, True), (3, 2, False), (2, 3, False), (3, 1, True)])
def test_oracle_connectivity_equals_reference_mesh_manager(cuda, dim, level, periodic, dtype):
    s = ref_cuda.RefSolver("uns", dtype, dim, level, periodic)
    f = oracle.Forest(dim, level, periodic)
    ref, orc = s.connectivity(), f.connectivity(dtype=dtype)
    assert_conn_equal(ref, orc)
    assert np.array_equal(ref["volumes"], f.elements()[2].astype(dtype))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oracle_adapt_equals_reference_adapt(cuda, dtype):
    """t8mini runs the reference's adapt callback + its level walk; the mini-forest restatement must produce the
    same forest, the same old->new map (observed through the remapped volumes) and the same connectivity."""
    s = ref_cuda.RefSolver("uns", dtype, 3, 3, True)
    f = oracle.Forest(3, 3, True)
    lv, cent, vol, _ = f.elements()
    for rnd in range(3):
        lv, cent, vol, _ = f.elements()
        rng = np.random.default_rng(rnd)
        crit = np.where(np.abs(cent[:, 2] - 0.5 + 0.1 * rnd) < 0.2, 20.0, 0.0) + rng.uniform(0, 1, len(lv))
        crit = crit.astype(dtype)
        s.mesh_adapt(crit)
        f = f.adapt(crit, 10.0, 1, 4)
        assert f.num_elements == s.counts()["n_local"]
        ref, orc = s.connectivity(), f.connectivity(dtype=dtype)
        assert_conn_equal(ref, orc)
        assert np.array_equal(ref["volumes"], f.elements()[2].astype(dtype))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,level,periodic", [(2, 5, True), (3, 3, True), (3, 3, False)])
def test_oracle_arithmetic_matches_reference_kernels(cuda, dim, level, periodic, dtype):
    s = ref_cuda.RefSolver("uns", dtype, dim, level, periodic)
    f = oracle.Forest(dim, level, periodic)
    conn = f.connectivity(dtype=dtype)
    u, vol = perturbed_kh(f, dtype, seed=21)
    s.set_state(u)
    dt = 0.1 * 2.0 ** -level
    for it in range(5):
        ref_prev = s.get_state()
        orc, _, _ = oracle.iterate(conn, vol, ref_prev, dt)     # restart from the reference's state: per-step error
        s.iterate(dt)
        err = rel_linf(orc, s.get_state())
        assert err <= TOL[np.dtype(dtype)], (it, err)
    # CFL reduction: reference compute_timestep (thrust::reduce + formula) vs oracle on the last stage's speeds
    sp = np.zeros(conn["n_faces"] + conn["n_bfaces"], dtype)
    oracle.iterate(conn, vol, ref_prev, dt, speed=sp)
    assert abs(s.compute_timestep() - oracle.compute_timestep(sp, dtype(0.7), 4)) <= 1e-5 * s.compute_timestep()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["fused", "unfused"])
@pytest.mark.parametrize("dim,level,periodic,adapt", [(2, 6, True, False), (3, 4, True, False), (3, 3, False, True)])
def test_product_matches_reference(cuda, dim, level, periodic, adapt, mode, dtype):
    """The product (C ABI) against the reference's own kernels, from the same state, per step and over 20 steps,
    on the arrays the REFERENCE's mesh manager built."""
    import t8gpu_b200
    s = ref_cuda.RefSolver("uns", dtype, dim, level, periodic)
    f = oracle.Forest(dim, level, periodic)
    if adapt:
        lv, cent, vol, _ = f.elements()
        crit = np.where(np.abs(cent[:, 2] - 0.5) < 0.2, 20.0, 0.0).astype(dtype)
        s.mesh_adapt(crit)
        f = f.adapt(crit, 10.0, 1, 4)
    conn = s.connectivity()
    conn["n_xfaces"] = 0
    u, vol = perturbed_kh(f, dtype, seed=22)
    s.set_state(u)
    sol = t8gpu_b200.EulerSolver(conn, conn["volumes"], DT[dtype], device=cuda, mode=mode)
    sol.set_state(u)
    dt = 0.1 * 2.0 ** -(level + (1 if adapt else 0))
    nsteps = 20
    for it in range(nsteps):
        s.iterate(dt)
        sol.iterate(dt)
        err = rel_linf(sol.state().cpu().numpy(), s.get_state())
        assert err <= (it + 1) * TOL[np.dtype(dtype)], (it, err)
    assert abs(sol.compute_timestep() - s.compute_timestep()) <= 1e-5 * s.compute_timestep()
