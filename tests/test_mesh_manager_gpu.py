"""t8gpu::MeshManager of the header mirror (include/t8gpu/mesh/mesh_manager.h), driven through tests/_headers/
mesh_harness.cu over the t8mini stand-in for t8code: connectivity arrays bit-exact vs the oracle and vs the reference's
own MeshManager, fused and reference-shaped stepping vs the oracle, adapt() vs the oracle's remap."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from oracle import ref_cuda
from util import TOL, perturbed_kh, rel_linf

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def _lib(prec):
    if prec not in _LIBS:
        so = os.path.join(HERE, "_headers", "libmeshharness_%s.so" % prec)
        if not os.path.exists(so):
            import sys
            sys.path.insert(0, os.path.join(HERE, "_headers"))
            import build as hb
            hb.build()
        L = C.CDLL(so)
        L.mh_create.restype = C.c_void_p
        L.mh_speed_max.restype = C.c_double
        assert L.mh_float_size() == (8 if prec == "f64" else 4)
        _LIBS[prec] = L
    return _LIBS[prec]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Mesh:
    def __init__(self, dtype, dim, level, periodic=True):
        self.dtype = np.dtype(dtype)
        self.L = _lib("f64" if self.dtype == np.float64 else "f32")
        self.h = C.c_void_p(self.L.mh_create(dim, level, int(periodic)))

    def close(self):
        if self.h:
            self.L.mh_destroy(self.h)
            self.h = None

    def counts(self):
        out = (C.c_int64 * 4)()
        self.L.mh_counts(self.h, out)
        return dict(n_local=out[0], n_ghost=out[1], n_faces=out[2], n_bfaces=out[3])

    def connectivity(self):
        c = self.counts()
        n, g, nf, nb = c["n_local"], c["n_ghost"], c["n_faces"], c["n_bfaces"]
        ranks, indices = np.zeros(n + g, np.int32), np.zeros(n + g, np.int32)
        nbr = np.zeros(2 * nf + nb, np.int32)
        normals, areas, vol = np.zeros(3 * (nf + nb), self.dtype), np.zeros(nf + nb, self.dtype), np.zeros(n, self.dtype)
        self.L.mh_get_connectivity(self.h, _p(ranks), _p(indices), _p(nbr), _p(normals), _p(areas), _p(vol))
        c.update(ranks=ranks, indices=indices, face_neighbors=nbr, face_normals=normals, face_areas=areas, volumes=vol)
        return c

    def set_state(self, u):
        u = np.ascontiguousarray(u, dtype=self.dtype)
        self.L.mh_set_state(self.h, _p(u))

    def get_state(self):
        u = np.zeros((5, self.counts()["n_local"]), self.dtype)
        self.L.mh_get_state(self.h, _p(u))
        return u

    def iterate(self, dt, n=1, fused=True):
        (self.L.mh_iterate if fused else self.L.mh_iterate_unfused)(self.h, C.c_double(dt), n)
        assert self.L.mh_last_cuda_error() == 0

    def criteria(self):
        out = np.zeros(self.counts()["n_local"], self.dtype)
        self.L.mh_criteria(self.h, _p(out))
        return out

    def adapt(self, crit):
        crit = np.ascontiguousarray(crit, dtype=self.dtype)
        self.L.mh_adapt(self.h, _p(crit))
        assert self.L.mh_last_cuda_error() == 0


KEYS = ("ranks", "indices", "face_neighbors", "face_normals", "face_areas")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,level,periodic", [(2, 6, True), (3, 3, True), (3, 3, False), (2, 4, False)])
def test_connectivity_bit_exact(cuda, dim, level, periodic, dtype):
    m = Mesh(dtype, dim, level, periodic)
    got = m.connectivity()
    f = oracle.Forest(dim, level, periodic)
    ref = f.connectivity(dtype=dtype)
    for k in ("n_local", "n_ghost", "n_faces", "n_bfaces"):
        assert got[k] == ref[k], k
    for k in KEYS:
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(got["volumes"], f.elements()[2].astype(dtype))
    m.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("fused", [True, False])
def test_step_adapt_step(cuda, dtype, fused):
    """The reference's AMR cycle on the product's manager: step, criteria, adapt (+partition + connectivity), step;
    checked against the oracle at every stage."""
    dim, level = 3, 3
    m = Mesh(dtype, dim, level, False)
    f = oracle.Forest(dim, level, False)
    u, vol = perturbed_kh(f, dtype, seed=12)
    m.set_state(u)
    dt = 0.05 * 2.0 ** -4
    conn = f.connectivity(dtype=dtype)
    for it in range(2):
        u, _, _ = oracle.iterate(conn, vol, u, dt)
    m.iterate(dt, 2, fused)
    assert rel_linf(m.get_state(), u) <= 2 * TOL[np.dtype(dtype)]
    u = m.get_state()
    # criteria of the example solver, then a criterion that refines a slab (so that the test does not depend on the
    # state crossing the threshold)
    crit = m.criteria()
    ref_crit = oracle.gradient_criteria(conn, u[0], vol)
    assert np.abs(crit - ref_crit).max() <= (1e-13 if dtype == np.float64 else 1e-5) * np.abs(ref_crit).max()
    lv, cent, _, _ = f.elements()
    crit = np.where(np.abs(cent[:, 2] - 0.5) < 0.2, 20.0, 0.0).astype(dtype)
    m.adapt(crit)
    f2 = f.adapt(crit, 10.0, 1, 4)
    u2, vol2 = oracle.adapt_remap(f.adapt_map(f2), u, vol, 0)
    got = m.connectivity()
    ref = f2.connectivity(dtype=dtype)
    for k in KEYS:
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(got["volumes"], vol2)
    assert np.array_equal(m.get_state(), u2)
    for it in range(2):
        u2, _, _ = oracle.iterate(ref, vol2, u2, dt)
    m.iterate(dt, 2, fused)
    assert rel_linf(m.get_state(), u2) <= 2 * TOL[np.dtype(dtype)]
    if fused:
        assert m.L.mh_speed_max(m.h) > 0
    m.close()


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_same_arrays_as_the_reference_manager(cuda, dtype):
    """Both managers over the same t8mini forest, after the same adapt: identical arrays."""
    s = ref_cuda.RefSolver("uns", dtype, 3, 3, False)
    m = Mesh(dtype, 3, 3, False)
    f = oracle.Forest(3, 3, False)
    lv, cent, _, _ = f.elements()
    crit = np.where(np.abs(cent[:, 0] - 0.5) < 0.2, 20.0, 0.0).astype(dtype)
    u0 = s.get_state()
    m.set_state(u0)
    s.mesh_adapt(crit)
    m.adapt(crit)
    a, b = s.connectivity(), m.connectivity()
    for k in KEYS + ("volumes",):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(s.get_state(), m.get_state())
    s.close()
    m.close()


# ------------------------------------------------------------------------------------------------ subgrid manager
_SLIBS = {}


def _slib(prec):
    if prec not in _SLIBS:
        L = C.CDLL(os.path.join(HERE, "_headers", "libsubgridharness_%s.so" % prec))
        L.sh_create.restype = C.c_void_p
        assert L.sh_float_size() == (8 if prec == "f64" else 4)
        _SLIBS[prec] = L
    return _SLIBS[prec]


class SubgridMesh:
    def __init__(self, dtype, dim, level, periodic=True):
        self.dtype, self.dim = np.dtype(dtype), dim
        self.S = 64 if dim == 3 else 16
        self.L = _slib("f64" if self.dtype == np.float64 else "f32")
        self.h = C.c_void_p(self.L.sh_create(dim, level, int(periodic)))

    def close(self):
        if self.h:
            self.L.sh_destroy(self.h)
            self.h = None

    def counts(self):
        out = (C.c_int64 * 4)()
        self.L.sh_counts(self.h, out)
        return dict(n_local=out[0], n_ghost=out[1], n_faces=out[2], n_bfaces=out[3])

    def connectivity(self):
        c = self.counts()
        n, g, nf, nb, d = c["n_local"], c["n_ghost"], c["n_faces"], c["n_bfaces"], self.dim
        ranks, indices = np.zeros(n + g, np.int32), np.zeros(n + g, np.int32)
        nbr = np.zeros(2 * nf + nb, np.int32)
        normals, areas, vol = np.zeros(d * (nf + nb), self.dtype), np.zeros(nf + nb, self.dtype), np.zeros(n, self.dtype)
        ld, off = np.zeros(nf, np.int32), np.zeros(d * nf, np.int32)
        self.L.sh_get_connectivity(self.h, _p(ranks), _p(indices), _p(nbr), _p(normals), _p(areas), _p(ld), _p(off), _p(vol))
        c.update(ranks=ranks, indices=indices, face_neighbors=nbr, face_normals=normals, face_areas=areas,
                 level_diff=ld, offsets=off, volumes=vol)
        return c

    def set_state(self, u):
        u = np.ascontiguousarray(u, dtype=self.dtype)
        self.L.sh_set_state(self.h, _p(u))

    def get_state(self):
        u = np.zeros((5, self.counts()["n_local"] * self.S), self.dtype)
        self.L.sh_get_state(self.h, _p(u))
        return u

    def iterate(self, dt, n=1):
        self.L.sh_iterate(self.h, C.c_double(dt), n)
        assert self.L.sh_last_cuda_error() == 0

    def criteria(self):
        out = np.zeros(self.counts()["n_local"], self.dtype)
        self.L.sh_criteria(self.h, _p(out))
        return out

    def adapt(self, crit):
        crit = np.ascontiguousarray(crit, dtype=self.dtype)
        self.L.sh_adapt(self.h, _p(crit))
        assert self.L.sh_last_cuda_error() == 0


SKEYS = KEYS + ("level_diff", "offsets")


def _sg_state(forest, dtype, seed):
    lv, cent, vol, _ = forest.elements()
    u = oracle.subgrid_init_kh(forest.dim, cent.astype(dtype), lv, dtype).astype(np.float64)
    rng = np.random.default_rng(seed)
    n = u.shape[1]
    rho = u[0] * (1 + 0.05 * rng.uniform(-1, 1, n))
    v = u[1:4] / u[0] + 0.05 * rng.uniform(-1, 1, (3, n))
    if forest.dim == 2:
        v[2] = 0.0
    p = 2.5 * (1 + 0.05 * rng.uniform(-1, 1, n))
    out = np.empty_like(u)
    out[0], out[1:4], out[4] = rho, rho * v, p / 0.4 + 0.5 * rho * (v * v).sum(0)
    return np.ascontiguousarray(out.astype(dtype)), vol.astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,level,periodic", [(3, 2, True), (2, 3, True), (3, 2, False), (2, 3, False)])
def test_subgrid_manager_amr_cycle(cuda, dim, level, periodic, dtype):
    """Connectivity bit-exact vs the oracle, fused stepping, criteria, adapt + partition + connectivity, stepping again."""
    m = SubgridMesh(dtype, dim, level, periodic)
    f = oracle.Forest(dim, level, periodic)
    got, ref = m.connectivity(), f.connectivity(subgrid=True, dtype=dtype)
    for k in SKEYS:
        assert np.array_equal(got[k], ref[k]), k
    u, vol = _sg_state(f, dtype, seed=21)
    assert np.array_equal(got["volumes"], vol)
    m.set_state(u)
    dt = 0.1 * 2.0 ** -(level + 3)
    for it in range(2):
        u, _, _ = oracle.subgrid_iterate(ref, vol, u, dt)
    m.iterate(dt, 2)
    assert rel_linf(m.get_state(), u) <= 2 * TOL[np.dtype(dtype)]
    u = m.get_state()
    crit = m.criteria()
    ref_crit = oracle.subgrid_criteria(dim, u[0], vol)
    assert np.abs(crit - ref_crit).max() <= (4e-13 if dtype == np.float64 else 4e-5) * np.abs(ref_crit).max()
    lv, cent, _, _ = f.elements()
    crit = np.where(np.abs(cent[:, dim - 1] - 0.5) < 0.2, 1.0, 0.0).astype(dtype)
    m.adapt(crit)
    f2 = f.adapt(crit, 0.02, 1, 6)
    u2, vol2 = oracle.adapt_remap(f.adapt_map(f2), u, vol, dim)
    got, ref = m.connectivity(), f2.connectivity(subgrid=True, dtype=dtype)
    for k in SKEYS:
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(got["volumes"], vol2)
    assert np.array_equal(m.get_state(), u2)
    for it in range(2):
        u2, _, _ = oracle.subgrid_iterate(ref, vol2, u2, dt / 2)
    m.iterate(dt / 2, 2)
    assert rel_linf(m.get_state(), u2) <= 2 * TOL[np.dtype(dtype)]
    m.close()


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,level", [(3, 2), (2, 3)])
def test_subgrid_same_arrays_as_the_reference_manager(cuda, dim, level, dtype):
    s = ref_cuda.RefSolver("sg", dtype, dim, level, True)
    m = SubgridMesh(dtype, dim, level, True)
    f = oracle.Forest(dim, level, True)
    lv, cent, _, _ = f.elements()
    crit = np.where(np.abs(cent[:, 0] - 0.5) < 0.2, 1.0, 0.0).astype(dtype)
    m.set_state(s.get_state())
    s.mesh_adapt(crit)
    m.adapt(crit)
    a, b = s.connectivity(), m.connectivity()
    for k in SKEYS + ("volumes",):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(s.get_state(), m.get_state())
    s.close()
    m.close()
