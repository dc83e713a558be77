"""TEST INFRASTRUCTURE ONLY: ctypes front-end of oracle/_ref -- the reference's OWN CUDA implementation of the hot path
(compiled unmodified from /root/reference by oracle/ref_build.py, linked against the t8mini shim).  Needs a GPU.

Used by the -m gpu parity tests, by tests/golden/make_golden.py (fixture generation) and by bench.py --impl reference.
"""
import ctypes as C
import json
import os
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
_LIBS = {}


def _path(kind, prec, mirror=False):
    return os.path.join(OUT, "lib%s_%s_%s.so" % ("mirror" if mirror else "ref", kind, prec))


def available():
    return all(os.path.exists(_path(k, p)) for k in ("uns", "sg") for p in ("f32", "f64"))


def mirror_available():
    """libmirror_*: the SAME unmodified reference example solvers + harness, compiled against the product's header
    mirror (include/t8gpu/) instead of the reference's t8gpu/ headers, linked with libt8gpu_b200.so."""
    return all(os.path.exists(_path(k, p, True)) for k in ("uns", "sg") for p in ("f32", "f64"))


def _lib(kind, prec, mirror=False):
    key = (kind, prec, mirror)
    if key not in _LIBS:
        L = C.CDLL(_path(kind, prec, mirror))
        L.t8mini_vtk_prefix.restype = C.c_char_p
        L.t8mini_vtk_field_name.restype = C.c_char_p
        L.t8mini_vtk_field.restype = C.c_int64
        L.ref_create.restype = C.c_void_p
        L.ref_time_steps.restype = C.c_double
        if kind == "uns":
            L.ref_compute_timestep.restype = C.c_double
        assert L.ref_float_size() == (8 if prec == "f64" else 4)
        _LIBS[key] = L
    return _LIBS[key]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefSolver:
    """kind="uns": t8gpu::CompressibleEulerSolver over MeshManager<VariableList,StepList,3>;
    kind="sg": SubgridCompressibleEulerSolver<Subgrid<4,4,4>> (dim 3) / <Subgrid<4,4>> (dim 2)."""

    def __init__(self, kind, dtype, dim, level, periodic=True, mirror=False):
        self.kind, self.dim = kind, dim
        self.dtype = np.dtype(dtype)
        self.prec = "f64" if self.dtype == np.float64 else "f32"
        self.L = _lib(kind, self.prec, mirror)
        self.h = C.c_void_p(self.L.ref_create(dim, level, int(periodic)))
        self.cells_per_element = (64 if dim == 3 else 16) if kind == "sg" else 1

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counts(self):
        out = (C.c_int64 * 4)()
        self.L.ref_counts(self.h, out)
        return dict(n_local=out[0], n_ghost=out[1], n_faces=out[2], n_bfaces=out[3])

    def connectivity(self):
        c = self.counts()
        n, g, nf, nb = c["n_local"], c["n_ghost"], c["n_faces"], c["n_bfaces"]
        nd = self.dim if self.kind == "sg" else 3
        ranks, indices = np.zeros(n + g, np.int32), np.zeros(n + g, np.int32)
        nbr = np.zeros(2 * nf + nb, np.int32)
        normals, areas = np.zeros(nd * (nf + nb), self.dtype), np.zeros(nf + nb, self.dtype)
        vol = np.zeros(n, self.dtype)
        c.update(dim=self.dim, ndim_normal=nd)
        if self.kind == "uns":
            self.L.ref_get_connectivity(self.h, _p(ranks), _p(indices), _p(nbr), _p(normals), _p(areas), _p(vol))
        else:
            ld, off = np.zeros(nf, np.int32), np.zeros(self.dim * nf, np.int32)
            self.L.ref_get_connectivity(self.h, _p(ranks), _p(indices), _p(nbr), _p(normals), _p(areas), _p(ld), _p(off),
                                        _p(vol))
            c.update(level_diff=ld, offsets=off)
        c.update(ranks=ranks, indices=indices, face_neighbors=nbr, face_normals=normals, face_areas=areas, volumes=vol)
        return c

    def num_cells(self):
        return self.counts()["n_local"] * self.cells_per_element

    def set_state(self, u):
        u = np.ascontiguousarray(u, dtype=self.dtype)
        assert u.shape == (5, self.num_cells())
        self.L.ref_set_state(self.h, _p(u))

    def get_state(self):
        u = np.zeros((5, self.num_cells()), self.dtype)
        self.L.ref_get_state(self.h, _p(u))
        return u

    def iterate(self, dt, nsteps=1):
        self.L.ref_iterate(self.h, C.c_double(dt), nsteps)
        err = self.L.ref_last_cuda_error()
        assert err == 0, "CUDA error %d in the reference" % err

    def adapt(self):
        self.L.ref_adapt(self.h)

    def mesh_adapt(self, crit):
        """MeshManager::adapt(criteria, next) [+ partition for the subgrid manager] + compute_connectivity_information."""
        crit = np.ascontiguousarray(crit, dtype=self.dtype)
        assert crit.shape == (self.counts()["n_local"],)
        self.L.ref_mesh_adapt(self.h, _p(crit))
        err = self.L.ref_last_cuda_error()
        assert err == 0, "CUDA error %d in the reference" % err

    def criteria(self):
        """The refinement criteria the reference's own adapt() would compute from the current state."""
        out = np.zeros(self.counts()["n_local"], self.dtype)
        self.L.ref_criteria(self.h, _p(out))
        err = self.L.ref_last_cuda_error()
        assert err == 0, "CUDA error %d in the reference" % err
        return out

    def compute_timestep(self):
        return self.L.ref_compute_timestep(self.h)

    def time_steps(self, dt, warmup, steps):
        return self.L.ref_time_steps(self.h, C.c_double(dt), warmup, steps)

    # ---- output path: what the solver's save_* members hand to t8_forest_write_vtk_ext (captured by t8mini)
    def save(self, what, prefix):
        """what: "conserved" (uns: save_conserved_variables_to_vtk), "density" / "mesh" (sg: save_density_to_vtk /
        save_mesh_to_vtk).  Returns the captured call: dict(calls, n_elements, dim, min_level, max_level, prefix,
        fields=[(name, array)])."""
        fn = {"conserved": "ref_save_conserved", "density": "ref_save_density", "mesh": "ref_save_mesh"}[what]
        getattr(self.L, fn)(self.h, prefix.encode())
        err = self.L.ref_last_cuda_error()
        assert err == 0, "CUDA error %d" % err
        info = (C.c_int64 * 6)()
        self.L.t8mini_vtk_info(info)
        fields = []
        for k in range(info[1]):
            n = self.L.t8mini_vtk_field(k, None, C.c_int64(0))
            a = np.zeros(n, np.float64)
            self.L.t8mini_vtk_field(k, _p(a), C.c_int64(n))
            fields.append((self.L.t8mini_vtk_field_name(k).decode(), a))
        return dict(calls=info[0], n_elements=info[2], dim=info[3], min_level=info[4], max_level=info[5],
                    prefix=self.L.t8mini_vtk_prefix().decode(), fields=fields)


def bench(args):
    """bench.py --impl reference: the reference's own iterate() (with its cudaDeviceSynchronize()s) on the same
    workload, same metric.  One rank only: the reference cannot place ranks on different GPUs (no cudaSetDevice, MPI
    + CUDA-IPC), and MPI is absent here."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    import oracle
    dtype = np.float64 if args.dtype == "f64" else np.float32
    level = args.level
    if getattr(args, "workload", "unstructured") == "subgrid":
        return bench_subgrid(args, dtype, level)
    t0 = time.time()
    s = RefSolver("uns", dtype, 3, level, True)   # host-side: t8mini forest + the reference's connectivity loop
    t_mesh = time.time() - t0
    n = s.counts()["n_local"]
    # same initial data as the product arm: Cartesian KH at the centroids
    f = oracle.Forest(3, level)
    lv, cent, vol, _ = f.elements()
    s.set_state(oracle.init_kh_points(3, cent.astype(dtype), dtype))
    dt = 0.1 * 2.0 ** -level
    ms = s.time_steps(dt, args.warmup, args.steps)
    u = s.get_state()
    assert np.isfinite(u).all()
    value = n * args.steps / (ms * 1e-3)
    line = {"impl": "reference", "metric": "cell-updates/s per RK3 step", "value": value, "unit": "cell-updates/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": "kelvin_helmholtz 3D uniform periodic hex mesh level %d (%d elements) %s, fixed dt, "
                                   "no adaptation" % (level, n, args.dtype),
                       "implementation": "reference CUDA kernels (examples/compressible_euler) compiled unmodified "
                                         "for sm_100a%s, reference iterate() schedule, 1 rank" %
                                         (" with float_type=double" if args.dtype == "f64" else ""),
                       "host_mesh_s": round(t_mesh, 3), "host_mesh": "t8mini (t8code absent), reference's own "
                       "compute_connectivity_information loop, 1 thread", "host_cores": os.cpu_count()},
            "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": 0, "kind": "reference",
                             "sample": "full workload on the GPU: the reference has no CPU implementation of this "
                                       "path; this is its own CUDA build on the same B200"},
            "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bench_subgrid(args, dtype, level):
    """--workload subgrid: SubgridCompressibleEulerSolver<Subgrid<4,4,4>>::iterate of the reference, one rank."""
    import oracle
    t0 = time.time()
    s = RefSolver("sg", dtype, 3, level, True)
    t_mesh = time.time() - t0
    f = oracle.Forest(3, level)
    lv, cent, vol, _ = f.elements()
    s.set_state(oracle.subgrid_init_kh(3, cent.astype(dtype), lv, dtype))
    n = s.num_cells()
    dt = 0.1 * 2.0 ** -(level + 2)
    ms = s.time_steps(dt, args.warmup, args.steps)
    assert np.isfinite(s.get_state()).all()
    value = n * args.steps / (ms * 1e-3)
    line = {"impl": "reference", "metric": "cell-updates/s per RK3 step", "value": value, "unit": "cell-updates/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": "kelvin_helmholtz 3D Subgrid<4,4,4> on a uniform periodic hex forest level %d (%d "
                                   "cells) %s, fixed dt, no adaptation" % (level, n, args.dtype),
                       "implementation": "reference CUDA kernels (examples/subgrid) compiled unmodified for sm_100a%s, "
                                         "reference iterate() schedule, 1 rank" %
                                         (" with float_type=double" if args.dtype == "f64" else ""),
                       "host_mesh_s": round(t_mesh, 3), "host_cores": os.cpu_count()},
            "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": 0, "kind": "reference",
                             "sample": "full workload on the GPU: the reference has no CPU implementation of this "
                                       "path; this is its own CUDA build on the same B200"},
            "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
