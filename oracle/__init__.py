"""TEST INFRASTRUCTURE ONLY: ctypes front-end of the CPU oracle (oracle/liboracle.so).

The oracle is a CPU restatement of the reference's hot path (see euler_impl.inc / miniforest.c for the reference
file:line each function follows).  It exists to CHECK the CUDA product; nothing under ``t8gpu_b200/`` imports it.
Allowed importers: ``tests/``, ``__graft_entry__.smoke()``, ``bench.py`` (cpu_baseline / reference legs).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("euler_oracle.c", "euler_impl.inc", "miniforest.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.mf_new_uniform.restype = C.c_void_p
        _LIB.mf_adapt.restype = C.c_void_p
        _LIB.mf_connectivity.restype = C.c_void_p
        _LIB.mf_num_elements.restype = C.c_int64
        for s in ("f32", "f64"):
            ft = C.c_float if s == "f32" else C.c_double
            getattr(_LIB, "orc_max_speed_" + s).restype = ft
            getattr(_LIB, "orc_compute_timestep_" + s).restype = ft
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", C.c_float
    if dtype == np.float64:
        return "f64", C.c_double
    raise TypeError(dtype)


class _Conn(C.Structure):
    _fields_ = [("n_local", C.c_int64), ("n_ghost", C.c_int64), ("n_faces", C.c_int64), ("n_bfaces", C.c_int64),
                ("ndim_normal", C.c_int), ("ranks", C.POINTER(C.c_int32)), ("indices", C.POINTER(C.c_int32)),
                ("ghost_global", C.POINTER(C.c_int64)), ("face_neighbors", C.POINTER(C.c_int32)),
                ("face_normals", C.POINTER(C.c_double)), ("face_areas", C.POINTER(C.c_double)),
                ("level_diff", C.POINTER(C.c_int32)), ("offsets", C.POINTER(C.c_int32)),
                ("n_xfaces", C.c_int64), ("x_face_neighbors", C.POINTER(C.c_int32)),
                ("x_face_normals", C.POINTER(C.c_double)), ("x_face_areas", C.POINTER(C.c_double)),
                ("x_level_diff", C.POINTER(C.c_int32)), ("x_offsets", C.POINTER(C.c_int32))]


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class Forest:
    """Cartesian one-tree quad/hex forest in Morton order (restated t8code semantics, miniforest.c)."""

    def __init__(self, dim, level, periodic=True, _handle=None):
        self.dim = dim
        self.periodic = periodic
        self._h = _handle if _handle is not None else lib().mf_new_uniform(dim, level, int(periodic))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mf_free(C.c_void_p(self._h))
            self._h = None

    @property
    def num_elements(self):
        return lib().mf_num_elements(C.c_void_p(self._h))

    def elements(self):
        n = self.num_elements
        levels = np.zeros(n, np.int32)
        cent = np.zeros((n, 3), np.float64)
        vol = np.zeros(n, np.float64)
        child = np.zeros(n, np.int32)
        lib().mf_get_elements(C.c_void_p(self._h), _p(levels), _p(cent), _p(vol), _p(child))
        return levels, cent, vol, child

    def partition_offsets(self, nranks):
        off = np.zeros(nranks + 1, np.int64)
        lib().mf_partition_offsets(C.c_void_p(self._h), nranks, _p(off))
        return off

    def adapt(self, crit, b, min_level, max_level, nranks=1):
        """t8_forest_set_adapt + balance with the reference's callback (mesh_manager.inl:125-162)."""
        crit = np.ascontiguousarray(crit)
        assert crit.shape == (self.num_elements,)
        is64 = crit.dtype == np.float64
        if not is64:
            crit = crit.astype(np.float32)
        off = self.partition_offsets(nranks)
        h = lib().mf_adapt(C.c_void_p(self._h), _p(crit), int(is64), C.c_double(b), min_level, max_level, _p(off),
                           nranks)
        return Forest(self.dim, 0, self.periodic, _handle=h)

    def adapt_map(self, new):
        """old->new element map (mesh_manager.inl:258-281)."""
        out = np.zeros(new.num_elements + 1, np.int32)
        lib().mf_adapt_map(C.c_void_p(self._h), C.c_void_p(new._h), 1 << self.dim, _p(out))
        return out

    def connectivity(self, nranks=1, rank=0, subgrid=False, ndim_normal=3, dtype=np.float64):
        """The arrays MeshManager / SubgridMeshManager::compute_connectivity_information would upload."""
        h = lib().mf_connectivity(C.c_void_p(self._h), nranks, rank, int(subgrid), ndim_normal)
        c = C.cast(h, C.POINTER(_Conn)).contents
        nl, ng, nf, nb, nd = c.n_local, c.n_ghost, c.n_faces, c.n_bfaces, c.ndim_normal
        out = dict(
            dim=self.dim, n_local=nl, n_ghost=ng, n_faces=nf, n_bfaces=nb, ndim_normal=nd, rank=rank, nranks=nranks,
            ranks=_arr(c.ranks, nl + ng, np.int32), indices=_arr(c.indices, nl + ng, np.int32),
            ghost_global=_arr(c.ghost_global, ng, np.int64),
            face_neighbors=_arr(c.face_neighbors, 2 * nf + nb, np.int32),
            face_normals=_arr(c.face_normals, nd * (nf + nb), dtype),
            face_areas=_arr(c.face_areas, nf + nb, dtype),
            n_xfaces=c.n_xfaces,
            x_face_neighbors=_arr(c.x_face_neighbors, 2 * c.n_xfaces, np.int32),
            x_face_normals=_arr(c.x_face_normals, nd * c.n_xfaces, dtype),
            x_face_areas=_arr(c.x_face_areas, c.n_xfaces, dtype),
        )
        if subgrid:
            out["level_diff"] = _arr(c.level_diff, nf, np.int32)
            out["offsets"] = _arr(c.offsets, self.dim * nf, np.int32)
            out["x_level_diff"] = _arr(c.x_level_diff, c.n_xfaces, np.int32)
            out["x_offsets"] = _arr(c.x_offsets, self.dim * c.n_xfaces, np.int32)
        lib().mf_conn_free(C.c_void_p(h))
        return out


# ---------------------------------------------------------------------------------------------- arithmetic

def face_flux(uL, uR, n, reflect=False):
    """Single-face flux in xyz (not area-scaled) and the wave-speed estimate."""
    uL = np.ascontiguousarray(uL)
    s, ft = _sfx(uL.dtype)
    uR = np.ascontiguousarray(uR, dtype=uL.dtype)
    n = np.ascontiguousarray(n, dtype=uL.dtype)
    F = np.zeros(5, uL.dtype)
    sp = ft(0)
    getattr(lib(), "orc_face_flux_" + s)(_p(uL), _p(uR), _p(n), int(reflect), _p(F), C.byref(sp))
    return F, sp.value


def flux_faces(conn, u, flux, speed=None, elem_index=None):
    """kepes_compute_fluxes + reflective_boundary_condition on SoA arrays u, flux of shape (5, stride)."""
    s, _ = _sfx(u.dtype)
    assert u.flags.c_contiguous and flux.flags.c_contiguous and u.shape == flux.shape
    getattr(lib(), "orc_flux_faces_" + s)(
        int(conn["n_faces"]), int(conn["n_bfaces"]), _p(conn["face_neighbors"]),
        _p(np.ascontiguousarray(conn["face_normals"], dtype=u.dtype)),
        _p(np.ascontiguousarray(conn["face_areas"], dtype=u.dtype)),
        _p(elem_index), _p(u), _p(flux), C.c_int64(u.shape[1]), _p(speed))


def rk_stage(stage, prev, inp, out, flux, vol, dt, cells_per_vol=1):
    s, ft = _sfx(prev.dtype)
    nvar, n = prev.shape
    getattr(lib(), "orc_rk_stage_" + s)(stage, C.c_int64(n), nvar, C.c_int64(n), _p(prev), _p(inp), _p(out),
                                        _p(flux), _p(np.ascontiguousarray(vol, dtype=prev.dtype)), cells_per_vol,
                                        ft(dt))


def iterate(conn, vol, prev, dt, speed=None):
    """One SSP-RK3 step of CompressibleEulerSolver::iterate on one rank; returns (next, step1, step2)."""
    s, ft = _sfx(prev.dtype)
    n = prev.shape[1]
    assert prev.shape[0] == 5 and prev.flags.c_contiguous
    s1, s2, nx, fl = (np.zeros_like(prev) for _ in range(4))
    if speed is None:
        speed = np.zeros(conn["n_faces"] + conn["n_bfaces"], prev.dtype)
    getattr(lib(), "orc_iterate_" + s)(
        C.c_int64(n), int(conn["n_faces"]), int(conn["n_bfaces"]), _p(conn["face_neighbors"]),
        _p(np.ascontiguousarray(conn["face_normals"], dtype=prev.dtype)),
        _p(np.ascontiguousarray(conn["face_areas"], dtype=prev.dtype)),
        _p(np.ascontiguousarray(vol, dtype=prev.dtype)), _p(prev), _p(s1), _p(s2), _p(nx), _p(fl), _p(speed), ft(dt))
    return nx, s1, s2


def max_speed(speed):
    s, _ = _sfx(speed.dtype)
    return getattr(lib(), "orc_max_speed_" + s)(_p(speed), C.c_int64(speed.size))


def compute_timestep(speed, cfl, max_level):
    s, ft = _sfx(speed.dtype)
    return getattr(lib(), "orc_compute_timestep_" + s)(_p(speed), C.c_int64(speed.size), ft(cfl), max_level)


def init_kh_points(dim, centers, dtype):
    """Cartesian Kelvin-Helmholtz field (subgrid/solver.inl:36-56) at points already cast to dtype -> (5, n)."""
    s, _ = _sfx(dtype)
    c = np.ascontiguousarray(centers, dtype=dtype)
    n = c.shape[0]
    u = np.zeros((5, n), dtype)
    getattr(lib(), "orc_init_kh_points_" + s)(dim, C.c_int64(n), _p(c), _p(u), C.c_int64(n))
    return u


def init_spherical_kh_points(centers, dtype):
    """Kelvin-Helmholtz field on the globe (compressible_euler/solver.cu:17-72) at points already cast to dtype."""
    s, _ = _sfx(dtype)
    c = np.ascontiguousarray(centers, dtype=dtype)
    n = c.shape[0]
    u = np.zeros((5, n), dtype)
    getattr(lib(), "orc_init_spherical_kh_points_" + s)(C.c_int64(n), _p(c), _p(u), C.c_int64(n))
    return u


def subgrid_init_kh(dim, centers, levels, dtype):
    s, _ = _sfx(dtype)
    c = np.ascontiguousarray(centers, dtype=dtype)
    ne = c.shape[0]
    S = 64 if dim == 3 else 16
    u = np.zeros((5, ne * S), dtype)
    getattr(lib(), "orc_subgrid_init_kh_" + s)(dim, C.c_int64(ne), _p(c), _p(np.ascontiguousarray(levels, np.int32)),
                                               _p(u), C.c_int64(ne * S))
    return u


def subgrid_iterate(conn, vol, prev, dt):
    """One SSP-RK3 step of SubgridCompressibleEulerSolver::iterate on one rank; returns (next, step1, step2)."""
    s, ft = _sfx(prev.dtype)
    dim = conn["dim"]
    S = 64 if dim == 3 else 16
    ne = prev.shape[1] // S
    s1, s2, nx, fl = (np.zeros_like(prev) for _ in range(4))
    getattr(lib(), "orc_subgrid_iterate_" + s)(
        dim, C.c_int64(ne), int(conn["n_faces"]), int(conn["n_bfaces"]), _p(conn["face_neighbors"]),
        _p(np.ascontiguousarray(conn["face_normals"], dtype=prev.dtype)),
        _p(np.ascontiguousarray(conn["face_areas"], dtype=prev.dtype)),
        _p(conn["level_diff"]), _p(conn["offsets"]), _p(np.ascontiguousarray(vol, dtype=prev.dtype)),
        _p(prev), _p(s1), _p(s2), _p(nx), _p(fl), ft(dt))
    return nx, s1, s2


def subgrid_flux(conn, vol, u, flux, elem_index=None):
    """inner + boundary + outer flux accumulation (one stage) on SoA cell arrays (5, ncells)."""
    s, _ = _sfx(u.dtype)
    dim = conn["dim"]
    S = 64 if dim == 3 else 16
    ne = len(vol)
    L = lib()
    nrm = np.ascontiguousarray(conn["face_normals"], dtype=u.dtype)
    ar = np.ascontiguousarray(conn["face_areas"], dtype=u.dtype)
    stride = C.c_int64(u.shape[1])
    getattr(L, "orc_subgrid_inner_" + s)(dim, C.c_int64(ne), _p(np.ascontiguousarray(vol, dtype=u.dtype)), _p(u),
                                         _p(flux), stride)
    if conn["n_bfaces"] > 0:
        getattr(L, "orc_subgrid_boundary_" + s)(dim, int(conn["n_faces"]), int(conn["n_bfaces"]),
                                                _p(conn["face_neighbors"]), _p(nrm), _p(ar), _p(u), _p(flux), stride)
    getattr(L, "orc_subgrid_outer_" + s)(dim, int(conn["n_faces"]), _p(conn["face_neighbors"]), _p(nrm), _p(ar),
                                         _p(conn["level_diff"]), _p(conn["offsets"]), _p(elem_index), _p(u), _p(flux),
                                         stride)


# ---------------------------------------------------------------------------------------------- adapt / partition remap

def adapt_remap(adapt_data, u_old, vol_old, subgrid_dim=0):
    """Restatement of adapt_variables_and_volume (t8gpu/mesh/mesh_manager.inl:164-193; subgrid_dim = 0) and of the
    subgrid manager's adapt_volume + adapt_variables (t8gpu/mesh/subgrid_mesh_manager.inl:245-425; subgrid_dim = 3 or 2).
    adapt_data: n_new + 1 ints (mesh_manager.inl:258-281); u_old: (nvar, n_old [* cells]); returns (u_new, vol_new) in
    u_old's dtype, with the reference's summation order."""
    ad = np.asarray(adapt_data, dtype=np.int64)
    dt = u_old.dtype
    n_new = len(ad) - 1
    diff = ad[1:] - ad[:-1]
    prev_same = np.zeros(n_new, bool)
    prev_same[1:] = ad[1:-1] == ad[:-2]
    if subgrid_dim == 2:
        fr, fc = dt.type(0.25), dt.type(4.0)
    else:
        fr, fc = dt.type(0.125), dt.type(8.0)     # also for MeshManager on 2-D meshes (SURVEY D-8)
    vol_new = vol_old[ad[:-1]] * np.where(diff == 0, fr, np.where(diff == 1, dt.type(1.0), fc)).astype(dt)
    vol_new[prev_same] = vol_old[ad[:-1]][prev_same] * fr
    vol_new = vol_new.astype(dt)
    nvar = u_old.shape[0]
    if subgrid_dim == 0:
        u_new = np.zeros((nvar, n_new), dt)
        for i in range(n_new):
            ns = max(1, int(diff[i]))
            for k in range(nvar):
                s = dt.type(0.0)
                for j in range(ns):
                    s = dt.type(s + dt.type(u_old[k, ad[i] + j] / dt.type(ns)))
                u_new[k, i] = s
        return u_new, vol_new
    dim = subgrid_dim
    S = 64 if dim == 3 else 16
    uo = u_old.reshape(nvar, -1, *([4] * dim)[::-1])          # [var][elem][(k)][j][i]
    un = np.zeros((nvar, n_new) + (4,) * dim, dt)
    for e in range(n_new):
        a = int(ad[e])
        if diff[e] == 0 or prev_same[e]:
            ri = 0
            while e - ri >= 0 and ad[e - ri] == a:
                ri += 1
            I, J, K = (ri - 1) & 1, ((ri - 1) >> 1) & 1, ((ri - 1) >> 2) & 1
            idx = np.arange(4) // 2
            if dim == 3:
                un[:, e] = uo[:, a][:, (K * 2 + idx)[:, None, None], (J * 2 + idx)[None, :, None], (I * 2 + idx)[None, None, :]]
            else:
                un[:, e] = uo[:, a][:, (J * 2 + idx)[:, None], (I * 2 + idx)[None, :]]
        elif diff[e] > 1:
            for c in range(S):
                i, j, k = c & 3, (c >> 2) & 3, (c >> 4) if dim == 3 else 0
                z = (i >> 1) | ((j >> 1) << 1) | ((k >> 1) << 2)
                for l in range(nvar):
                    s = dt.type(0.0)
                    for ii in range(2):
                        for jj in range(2):
                            for kk in range(2 if dim == 3 else 1):
                                if dim == 3:
                                    s = dt.type(s + uo[l, a + z, 2 * (k & 1) + kk, 2 * (j & 1) + jj, 2 * (i & 1) + ii])
                                else:
                                    s = dt.type(s + uo[l, a + z, 2 * (j & 1) + jj, 2 * (i & 1) + ii])
                    v = dt.type(s / dt.type(1 << dim))
                    if dim == 3:
                        un[l, e, k, j, i] = v
                    else:
                        un[l, e, j, i] = v
        else:
            un[:, e] = uo[:, a]
    return un.reshape(nvar, n_new * S), vol_new


def partition_remap(ranks, indices, u_old_by_rank, vol_old_by_rank, cells_per_element=1):
    """Restatement of partition_data (mesh_manager.inl:625-643) / partition_variable_data + partition_volume_data
    (subgrid_mesh_manager.inl:1216-1283): new element e <- old element indices[e] of rank ranks[e]."""
    S = cells_per_element
    nvar = u_old_by_rank[0].shape[0]
    n = len(ranks)
    dt = u_old_by_rank[0].dtype
    u = np.zeros((nvar, n * S), dt)
    vol = np.zeros(n, dt)
    for e in range(n):
        r, i = int(ranks[e]), int(indices[e])
        u[:, e * S:(e + 1) * S] = u_old_by_rank[r][:, i * S:(i + 1) * S]
        vol[e] = vol_old_by_rank[r][i]
    return u, vol


# ---------------------------------------------------------------------------------------------- refinement criteria

def gradient_criteria(conn, rho, vol):
    """Restatement of estimate_gradient + compute_refinement_criteria of the unstructured example
    (examples/compressible_euler/kernels.cu:471-501, solver.cu:231-245), single rank: every INTERIOR face adds
    |rho_R - rho_L| to both of its elements; criteria = sum / cbrt(volume).  (The reference accumulates with atomics, so
    its summation order is not defined; this one adds in face order.)"""
    dt = rho.dtype
    nf = int(conn["n_faces"])
    nbr = conn["face_neighbors"][:2 * nf].reshape(-1, 2)
    g = np.abs(rho[nbr[:, 1]] - rho[nbr[:, 0]]).astype(dt)
    acc = np.zeros(len(rho), dt)
    np.add.at(acc, nbr[:, 0], g)
    np.add.at(acc, nbr[:, 1], g)
    return (acc / np.cbrt(vol.astype(dt))).astype(dt)


def subgrid_criteria(dim, rho, vol):
    """Restatement of compute_refinement_criteria<Subgrid> (examples/subgrid/kernels.inl:1109-1168): H1 seminorm of
    the density over the cells of each element, summed in the reference's loop order, divided by the volume."""
    dt = rho.dtype
    n = len(vol)
    d = rho.reshape((n,) + (4,) * dim)          # [e][(r)][q][p], p fastest
    at = (lambda p, q, r: d[:, r, q, p]) if dim == 3 else (lambda p, q, r: d[:, q, p])
    h = (np.cbrt(vol) if dim == 3 else np.sqrt(vol)).astype(dt) / dt.type(4)
    s = np.zeros(n, dt)
    R = range(4) if dim == 3 else range(1)
    for p in range(3):
        for q in range(4):
            for r in R:
                x = at(p + 1, q, r) - at(p, q, r)
                s = (s + x * x * h).astype(dt)
    for p in range(4):
        for q in range(3):
            for r in R:
                x = at(p, q + 1, r) - at(p, q, r)
                s = (s + x * x * h).astype(dt)
    if dim == 3:
        for p in range(4):
            for q in range(4):
                for r in range(3):
                    x = at(p, q, r + 1) - at(p, q, r)
                    s = (s + x * x * h).astype(dt)
    return (s / vol.astype(dt)).astype(dt)
