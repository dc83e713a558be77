"""CPU suite: pins the oracle with analytic known-answer tests (the reference ships no tests or golden vectors,
SURVEY.md section 4) and against the committed golden fixtures generated from the reference's own CUDA kernels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def test_ln_mean_limit_and_freestream():
    # uniform state: flux must equal the physical Euler flux, dissipation vanishes (vJump == 0)
    rho, m, E = 1.3, np.array([0.3, 0.1, -0.2]), 2.5
    u = np.array([rho, *m, E])
    for n in (np.array([1.0, 0, 0]), np.array([0, -1.0, 0]), np.array([0.6, 0.0, 0.8])):
        F, s = oracle.face_flux(u, u, n)
        v = m / rho
        p = 0.4 * (E - 0.5 * rho * v @ v)
        vn = v @ n
        Fx = np.array([rho * vn, *(m * vn + p * n), (E + p) * vn])
        assert np.allclose(F, Fx, rtol=1e-13, atol=1e-14)
        assert np.isclose(s, abs(vn) + np.sqrt(1.4 * p / rho), rtol=1e-13)


def test_wall_bc_zero_mass_and_energy_flux():
    u = np.array([1.7, 0.4, -0.3, 0.2, 3.0])
    n = np.array([0.0, 0.6, 0.8])
    F, _ = oracle.face_flux(u, u, n, reflect=True)
    assert abs(F[0]) < 1e-15 and abs(F[4]) < 1e-14
    # momentum flux is along n
    t = np.cross(n, [1.0, 0, 0])
    assert abs(F[1:4] @ t) < 1e-14


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,level", [(2, 4), (3, 3)])
def test_conservation_and_freestream(dim, level, dtype):
    from util import perturbed_kh
    f = oracle.Forest(dim, level)
    conn = f.connectivity(dtype=dtype)
    u, vol = perturbed_kh(f, dtype)
    nx, s1, s2 = oracle.iterate(conn, vol, u, 0.1 * 2.0 ** -level)
    eps = np.finfo(dtype).eps
    for k in range(5):
        tot0, tot1 = (vol.astype(np.float64) * u[k]).sum(), (vol.astype(np.float64) * nx[k]).sum()
        assert abs(tot1 - tot0) <= 200 * eps * max(1.0, abs(tot0))
    # free stream is preserved exactly up to rounding
    uc = np.tile(np.array([[1.0], [0.3], [0.2], [0.0 if dim == 2 else -0.1], [2.5]], dtype=dtype), (1, u.shape[1]))
    nx, _, _ = oracle.iterate(conn, vol, np.ascontiguousarray(uc), 0.1 * 2.0 ** -level)
    # ... except that the reference's truncated stage-3 coefficients sum to 0.99999999999999, not 1
    # (ssp_runge_kutta.inl:23-25): a constant state decays by that factor each step.
    csum = dtype(0.33333333333333) + dtype(0.66666666666666)
    assert np.abs(nx - uc * csum).max() <= 50 * eps
    if dtype == np.float64:
        assert np.abs(nx[0] - 1.0).max() > 5e-15


def test_miniforest_uniform_counts_and_order():
    for dim, level in ((2, 6), (3, 3)):
        f = oracle.Forest(dim, level)
        c = f.connectivity()
        n = f.num_elements
        assert c["n_faces"] == dim * n and c["n_bfaces"] == 0 and c["n_ghost"] == 0
        nbr = c["face_neighbors"].reshape(-1, 2)
        assert (nbr[:, 0] < nbr[:, 1]).all()           # emitted by the lower element (mesh_manager.inl:411-414)
        assert (np.diff(nbr[:, 0]) >= 0).all()         # element-major order
        nrm = c["face_normals"].reshape(-1, 3)
        assert set(np.unique(np.abs(nrm))) <= {0.0, 1.0}
        assert np.all(c["face_areas"] == 2.0 ** (-(dim - 1) * level))


def test_miniforest_nonperiodic_boundary_faces():
    f = oracle.Forest(3, 2, periodic=False)
    c = f.connectivity()
    assert c["n_bfaces"] == 6 * 16 and c["n_faces"] == 3 * 4 * 4 * 3
    assert len(c["face_neighbors"]) == 2 * c["n_faces"] + c["n_bfaces"]


def test_miniforest_adapt_balance_partition():
    f = oracle.Forest(3, 3)
    lv, cent, vol, _ = f.elements()
    crit = np.where(np.abs(cent[:, 2] - 0.5) < 0.2, 20.0, 0.0)
    g = f.adapt(crit, 10.0, 1, 4)
    lv2, c2, vol2, _ = g.elements()
    assert np.isclose(vol2.sum(), 1.0)
    amap = f.adapt_map(g)
    assert amap[-1] == f.num_elements and (np.diff(amap) >= 0).all()
    cg = g.connectivity()
    # 2:1 balance: hanging faces have area ratio exactly 1/4 of the coarse face
    nbr = cg["face_neighbors"].reshape(-1, 2)
    assert np.abs(lv2[nbr[:, 0]] - lv2[nbr[:, 1]]).max() <= 1
    # every face is owned exactly once across ranks; x-faces are the complement on the higher rank
    for P in (2, 3, 5):
        tot = sum(g.connectivity(P, r)["n_faces"] for r in range(P))
        assert tot == cg["n_faces"]
        totx = sum(g.connectivity(P, r)["n_xfaces"] for r in range(P))
        ghostfaces = 0
        for r in range(P):
            cr = g.connectivity(P, r)
            ghostfaces += (cr["face_neighbors"][:2 * cr["n_faces"]].reshape(-1, 2) >= cr["n_local"]).any(1).sum()
        assert totx == ghostfaces


@pytest.mark.parametrize("dim", [2, 3])
def test_subgrid_matches_fine_uniform_grid(dim):
    """A uniform level-L forest of Subgrid<4,...> elements is the same discretisation as a uniform level-(L+2)
    forest of plain elements: the two oracle paths (kernels.inl vs kernels.cu restatements) must agree."""
    L = 2
    fs, ff = oracle.Forest(dim, L), oracle.Forest(dim, L + 2)
    cs = fs.connectivity(subgrid=True)
    cf = ff.connectivity()
    lvs, cents, vols, _ = fs.elements()
    lvf, centf, volf, _ = ff.elements()
    us = oracle.subgrid_init_kh(dim, cents, lvs, np.float64)
    uf = oracle.init_kh_points(dim, centf, np.float64)
    # map: cell centre -> fine element
    S = 4 ** dim
    h = 2.0 ** -(L + 2)
    key = {}
    for i, c in enumerate(centf):
        key[tuple(np.round(c[:dim] / h - 0.5).astype(int))] = i
    perm = np.zeros(fs.num_elements * S, dtype=np.int64)
    for e in range(fs.num_elements):
        for k in range(4 if dim == 3 else 1):
            for j in range(4):
                for i in range(4):
                    ijk = (i, j, k)
                    cc = [cents[e][d] - 0.5 * 2.0 ** -L + (ijk[d] + 0.5) * h for d in range(dim)]
                    perm[e * S + i + 4 * j + 16 * k] = key[tuple(np.round(np.array(cc) / h - 0.5).astype(int))]
    assert np.array_equal(us, uf[:, perm])
    dt = 0.1 * h
    ns, _, _ = oracle.subgrid_iterate(cs, vols, us, dt)
    nf, _, _ = oracle.iterate(cf, volf, uf, dt)
    assert np.abs(ns - nf[:, perm]).max() < 5e-14


def test_product_flux_algebra_on_host_matches_oracle():
    """The product's rotation-free flux (t8gpu_b200/csrc/euler_flux.cuh) compiled for the host must agree with the
    oracle's literal restatement of the reference arithmetic to rounding."""
    d = os.path.join(HERE, "_hostcheck")
    so = os.path.join(d, "libhostcheck.so")
    src = os.path.join(d, "host_flux_check.cu")
    hdr = os.path.join(HERE, "..", "t8gpu_b200", "csrc", "euler_flux.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
                               "-o", so, src])
    L = C.CDLL(so)
    L.hostcheck_flux_f64.restype = C.c_double
    L.hostcheck_flux_f32.restype = C.c_float
    rng = np.random.default_rng(0)

    def p(a):
        return a.ctypes.data_as(C.c_void_p)

    worst = {np.float64: 0.0, np.float32: 0.0}
    for it in range(4000):
        rho = rng.uniform(0.5, 2.5, 2)
        if it % 3 == 0:
            rho[1] = rho[0] * (1 + rng.uniform(-1e-3, 1e-3))
        v = rng.uniform(-1, 1, (2, 3))
        pr = rng.uniform(0.5, 3, 2)
        if it % 5 == 0:
            pr[1] = pr[0] * (1 + rng.uniform(-1e-4, 1e-4))
        E = pr / 0.4 + 0.5 * rho * (v * v).sum(1)
        uL = np.array([rho[0], *(rho[0] * v[0]), E[0]])
        uR = np.array([rho[1], *(rho[1] * v[1]), E[1]])
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        if it % 4 == 0:
            n = np.eye(3)[it % 3] * (1 if it % 8 else -1)
        refl = it % 7 == 0
        if refl:
            uR = uL.copy()
        scale = pr.max() + (rho * ((v * v).sum(1) + np.sqrt(1.4 * pr / rho) * np.abs(v).max(1))).max()
        for dt, fn in ((np.float64, L.hostcheck_flux_f64), (np.float32, L.hostcheck_flux_f32)):
            a, b, c = uL.astype(dt), uR.astype(dt), n.astype(dt)
            Fo, so_ = oracle.face_flux(a, b, c, refl)
            F = np.zeros(5, dt)
            s = fn(p(a), p(b), p(c), int(refl), p(F))
            worst[dt] = max(worst[dt], np.abs(F.astype(np.float64) - Fo).max() / scale, abs(s - so_) / so_)
    assert worst[np.float64] < 2e-14, worst
    assert worst[np.float32] < 2e-5, worst


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_spherical_kh_restatement_properties(dtype):
    """compressible_euler/solver.cu:17-72 restated: density 2 below / 1 above the equator, velocity tangent to the globe,
    zonal speed 0.5 r cos(theta) away from the perturbation, energy = 2.5 / 0.4 + kinetic part (pinned against the
    reference's own constructor in tests/test_initial_conditions_gpu.py)."""
    rng = np.random.default_rng(1)
    c = rng.uniform(-1.0, 1.0, (4000, 3))
    c *= (0.6 + 0.4 * rng.random((4000, 1))) / np.linalg.norm(c, axis=1, keepdims=True)
    c = c.astype(dtype)
    u = oracle.init_spherical_kh_points(c, dtype).astype(np.float64)
    cd = c.astype(np.float64)
    r = np.linalg.norm(cd, axis=1)
    theta = np.arcsin(cd[:, 2] / r)
    eps = np.finfo(dtype).eps
    assert np.array_equal(u[0], np.where(theta < 0, 2.0, 1.0))
    v = u[1:4] / u[0]
    assert np.abs((v * (cd / r[:, None]).T).sum(0)).max() < 50 * eps
    far = np.abs(theta) > 1.2                                    # exp(-(theta / 0.283)^2) < 2e-8 there
    speed = np.linalg.norm(v, axis=0)
    assert np.abs(speed[far] - 0.5 * r[far] * np.cos(theta[far])).max() < 1e-6
    assert np.abs(u[4] - (2.5 / 0.4 + 0.5 * (u[1:4] ** 2).sum(0) / u[0])).max() < 50 * eps * u[4].max()
