"""t8gpu_b200 -- B200-native finite-volume solver core behind t8gpu's hot path.

The product is the CUDA library ``libt8gpu_b200.so`` (sources in ``csrc/``, C ABI in ``include/t8gpu_b200.h``) and the
header-only C++ mirror of the reference's template API in ``include/t8gpu/``.  This Python package is harness glue for
tests and ``bench.py``: it binds the C ABI with ctypes and uses torch only for device memory and streams.  There is no
CPU fallback: importing works without a GPU (so that the symbol table can be checked), but every compute entry point
needs the library and a CUDA device and raises otherwise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# T8GPU_B200_LIB: alternative build of the same ABI (A/B timing of kernel variants in one gpurun call)
LIB_PATH = os.environ.get("T8GPU_B200_LIB") or os.path.join(_HERE, "libt8gpu_b200.so")
_LIB = None

NVAR = 5


class MissingExtension(RuntimeError):
    pass


def lib():
    """The C-ABI library.  Raises loudly if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MissingExtension(
                "t8gpu_b200/libt8gpu_b200.so is missing: run `python -m t8gpu_b200.build` (needs nvcc, sm_100a). "
                "There is no CPU fallback for this path.")
        _LIB = C.CDLL(LIB_PATH)
        _LIB.t8b200_plan_destroy.restype = None
        _LIB.t8b200_plan_ghost_tail_count.restype = C.c_int64
        _LIB.t8b200_plan_device_array.restype = C.c_int64
        _LIB.t8b200_cartesian_connectivity_free.restype = None
        if hasattr(_LIB, "t8b200_solver_destroy"):
            _LIB.t8b200_solver_destroy.restype = None
        if hasattr(_LIB, "t8b200_subgrid_plan_destroy"):
            _LIB.t8b200_subgrid_plan_destroy.restype = None
    return _LIB


class CudaError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        raise CudaError("t8gpu_b200: %s failed with cudaError %d" % (what, rc))


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise MissingExtension("t8gpu_b200 needs a CUDA device; there is no CPU fallback for this path.")
    return torch


def _sfx(dtype):
    import torch
    if dtype in (torch.float32, "f32"):
        return "f32", C.c_float
    if dtype in (torch.float64, "f64"):
        return "f64", C.c_double
    raise TypeError("float_type must be float32 or float64, got %r" % (dtype,))


def ptrs(tensors):
    """Host array of device pointers (one per variable), as MemoryAccessorOwn holds them."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def stream_ptr(stream=None):
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class RankTables:
    """Device tables `[var][rank] -> pointer` (what MemoryAccessorAll<VariableList> holds, memory_manager.h:300).

    `per_rank_vars`: list over ranks of lists of NVAR tensors (peer-accessible)."""

    def __init__(self, per_rank_vars, device):
        torch = _torch()
        nranks = len(per_rank_vars)
        host = torch.empty((NVAR, nranks), dtype=torch.int64)
        for r, vs in enumerate(per_rank_vars):
            for k in range(NVAR):
                host[k, r] = vs[k].data_ptr()
        self.table = host.to(device)
        self._keep = per_rank_vars
        self.host = (C.c_void_p * NVAR)()
        for k in range(NVAR):
            self.host[k] = self.table[k].data_ptr()


# ------------------------------------------------------------------------------------------------ reference-shaped

def flux_faces(conn, vars_all, flux_all, speed=None, stream=None):
    """kepes_compute_fluxes + reflective_boundary_condition.  conn: dict of device tensors in the reference layout
    (ranks, indices may be None for one rank); vars_all / flux_all: RankTables."""
    normals = conn["face_normals"]
    s, _ = _sfx(normals.dtype)
    fn = getattr(lib(), "t8b200_flux_faces_" + s)
    rk, ix = conn.get("ranks"), conn.get("indices")
    check(fn(int(conn["n_faces"]), int(conn["n_bfaces"]),
             C.c_void_p(rk.data_ptr() if rk is not None else None),
             C.c_void_p(ix.data_ptr() if ix is not None else None),
             C.c_void_p(conn["face_neighbors"].data_ptr()), C.c_void_p(normals.data_ptr()),
             C.c_void_p(conn["face_areas"].data_ptr()), vars_all.host, flux_all.host,
             C.c_void_p(speed.data_ptr() if speed is not None else None), stream_ptr(stream)), "flux_faces")


def rk3_stage(stage, prev, inp, out, flux, vol, dt, cells_per_vol=1, stream=None):
    """SSP_3RK_step{stage}; prev/inp/out/flux: lists of per-variable device tensors."""
    s, ft = _sfx(prev[0].dtype)
    fn = getattr(lib(), "t8b200_rk3_stage_" + s)
    n = prev[0].numel()
    check(fn(stage, C.c_int64(n), len(prev), ptrs(prev), ptrs(inp) if inp is not None else None, ptrs(out),
             ptrs(flux), C.c_void_p(vol.data_ptr()), cells_per_vol, ft(dt), stream_ptr(stream)), "rk3_stage")


def max_speed(speed, out=None, stream=None):
    torch = _torch()
    s, _ = _sfx(speed.dtype)
    if out is None:
        out = torch.zeros(1, dtype=speed.dtype, device=speed.device)
    check(getattr(lib(), "t8b200_max_speed_" + s)(C.c_void_p(speed.data_ptr()), C.c_int64(speed.numel()),
                                                   C.c_void_p(out.data_ptr()), stream_ptr(stream)), "max_speed")
    return out


def subgrid_inner_flux(dim, vol, vars_own, flux_own, stream=None):
    """compute_inner_fluxes<Subgrid<4,4[,4]>>; vars_own / flux_own: lists of 5 per-variable cell tensors."""
    s, _ = _sfx(vol.dtype)
    check(getattr(lib(), "t8b200_subgrid_inner_flux_" + s)(dim, C.c_int64(vol.numel()), C.c_void_p(vol.data_ptr()),
                                                            ptrs(vars_own), ptrs(flux_own), stream_ptr(stream)),
          "subgrid_inner_flux")


def _opt(t):
    return C.c_void_p(t.data_ptr() if t is not None else None)


def subgrid_outer_flux(conn, vars_all, flux_all, stream=None):
    """compute_outer_fluxes; conn: dict of device tensors in the SubgridMeshConnectivityAccessor layout."""
    s, _ = _sfx(conn["face_normals"].dtype)
    check(getattr(lib(), "t8b200_subgrid_outer_flux_" + s)(
        int(conn["dim"]), int(conn["n_faces"]), _opt(conn.get("ranks")), _opt(conn.get("indices")),
        _opt(conn["face_neighbors"]), _opt(conn["face_normals"]), _opt(conn["face_areas"]), _opt(conn["level_diff"]),
        _opt(conn["offsets"]), vars_all.host, flux_all.host, stream_ptr(stream)), "subgrid_outer_flux")


def subgrid_boundary_flux(conn, vars_own_tab, flux_own_tab, stream=None):
    """compute_boundary_fluxes on the n_bfaces wall faces; *_tab: RankTables of this rank only."""
    s, _ = _sfx(conn["face_normals"].dtype)
    check(getattr(lib(), "t8b200_subgrid_boundary_flux_" + s)(
        int(conn["dim"]), int(conn["n_faces"]), int(conn["n_bfaces"]), _opt(conn["face_neighbors"]),
        _opt(conn["face_normals"]), _opt(conn["face_areas"]), vars_own_tab.host, flux_own_tab.host,
        stream_ptr(stream)), "subgrid_boundary_flux")


# ------------------------------------------------------------------------------------------------ fused tile plan

class Plan:
    """Tile plan: the reference-layout connectivity re-laid out per chunk of elements."""

    def __init__(self, conn, dtype, ghost_tail=False):
        """conn: dict of HOST numpy arrays in the reference layout (as MeshManager::compute_connectivity_information
        builds them): n_local, n_ghost, n_faces, n_bfaces, face_neighbors, face_normals, face_areas, and for
        multi-rank ranks, indices, optionally x_face_* (faces whose ghost neighbour belongs to a lower rank).
        ghost_tail: the ghosts get local copies behind this rank's own elements (rows need n_local + n_tail entries),
        filled by pull() before each stage; the stage kernels then read no peer memory."""
        import numpy as np
        _torch()
        s, _ = _sfx(dtype)
        npdt = np.float64 if s == "f64" else np.float32
        self.dtype = dtype
        nrm = np.ascontiguousarray(conn["face_normals"], dtype=npdt)
        ar = np.ascontiguousarray(conn["face_areas"], dtype=npdt)
        nbr = np.ascontiguousarray(conn["face_neighbors"], dtype=np.int32)
        rk = conn.get("ranks")
        ix = conn.get("indices")
        rk = None if rk is None else np.ascontiguousarray(rk, dtype=np.int32)
        ix = None if ix is None else np.ascontiguousarray(ix, dtype=np.int32)
        nx = int(conn.get("n_xfaces", 0))
        xn = np.ascontiguousarray(conn["x_face_neighbors"], dtype=np.int32) if nx else None
        xr = np.ascontiguousarray(conn["x_face_normals"], dtype=npdt) if nx else None
        xa = np.ascontiguousarray(conn["x_face_areas"], dtype=npdt) if nx else None

        def p(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)

        h = C.c_void_p()
        create = lib().t8b200_plan_create_ghost_tail if ghost_tail else lib().t8b200_plan_create
        check(create(C.byref(h), int(s == "f64"), C.c_int64(int(conn["n_local"])),
                     C.c_int64(int(conn.get("n_ghost", 0))), int(conn["n_faces"]),
                     int(conn["n_bfaces"]), p(nbr), p(nrm), p(ar), p(rk), p(ix), nx, p(xn), p(xr),
                     p(xa)), "plan_create")
        self._h = h
        self.n_tail = int(lib().t8b200_plan_ghost_tail_count(h))
        info = (C.c_int64 * 8)()
        check(lib().t8b200_plan_info(self._h, info), "plan_info")
        self.info = dict(n_chunks=info[0], max_halo=info[1], max_faces=info[2], smem_bytes=info[3],
                         device_bytes=info[4], face_records=info[5], halo_entries=info[6], chunk=info[7])

    @classmethod
    def from_device(cls, conn, dtype, ghost_tail=False, stream=None):
        """Plan built on the device from DEVICE connectivity tensors (t8b200_plan_create_device): no copy of the
        connectivity to the host: the three-kernel builder for structured-only meshes, else one CUDA thread per block
        of 256 elements (csrc/plan_block.cuh).  Returns None only for a rank without elements (the caller then uses
        Plan(conn_to_host(conn), ...))."""
        _torch()
        s, _ = _sfx(dtype)

        def p(k):
            t = conn.get(k)
            return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())

        nx = int(conn.get("n_xfaces", 0))
        ng = int(conn.get("n_ghost", 0))
        h = C.c_void_p()
        rc = lib().t8b200_plan_create_device(
            C.byref(h), int(s == "f64"), int(bool(ghost_tail)), C.c_int64(int(conn["n_local"])), C.c_int64(ng),
            int(conn["n_faces"]), int(conn["n_bfaces"]), p("face_neighbors"), p("face_normals"), p("face_areas"),
            p("ranks") if ng else None, p("indices") if ng else None, nx, p("x_face_neighbors") if nx else None,
            p("x_face_normals") if nx else None, p("x_face_areas") if nx else None, stream_ptr(stream))
        if rc == 801:   # cudaErrorNotSupported: nothing to build
            return None
        check(rc, "plan_create_device")
        self = cls.__new__(cls)
        self.dtype, self._h = dtype, h
        self.n_tail = int(lib().t8b200_plan_ghost_tail_count(h))
        info = (C.c_int64 * 8)()
        check(lib().t8b200_plan_info(self._h, info), "plan_info")
        self.info = dict(n_chunks=info[0], max_halo=info[1], max_faces=info[2], smem_bytes=info[3],
                         device_bytes=info[4], face_records=info[5], halo_entries=info[6], chunk=info[7],
                         built_on="device")
        return self

    def device_array(self, which):
        """Structured / ghost-tail arrays of the plan copied from the device (13 s_rec, 14 s_halo, 15 s_hrank, 17
        pull_rank, 18 pull_idx) -> numpy int32."""
        import numpy as np
        n = lib().t8b200_plan_device_array(self._h, which, None, C.c_int64(0))
        if n < 0:
            raise CudaError("plan_device_array(%d)" % which)
        a = np.zeros(n, np.int32)
        if n:
            lib().t8b200_plan_device_array(self._h, which, a.ctypes.data_as(C.c_void_p), C.c_int64(n))
        return a

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:   # `lib` is gone when the interpreter shuts down
            lib().t8b200_plan_destroy(self._h)
            self._h = None

    def pull(self, rows, rows_all, stream=None):
        """Ghost tail <- the owners' rows (t8b200_ghost_pull).  rows: this rank's per-variable tensors of one step (with
        the tail behind the n_local own entries); rows_all: RankTables / PointerTables of the same step."""
        s, _ = _sfx(self.dtype)
        check(getattr(lib(), "t8b200_ghost_pull_" + s)(self._h, len(rows), ptrs(rows), rows_all.host,
                                                       stream_ptr(stream)), "ghost_pull")

    def stage_push(self, stage, inp, prev, out, out_all, vol, dt, send_csr, speed_max=None, stream=None, dt_dev=None):
        """A stage with the ghost push folded into the kernel (t8b200_fused_stage_push): out_all = tables of the output
        step, send_csr = (send_off, send_rank, send_idx) device int32 tensors (multi.send_csr).  False: not supported."""
        s, ft = _sfx(self.dtype)
        rc = getattr(lib(), "t8b200_fused_stage_push_" + s)(
            self._h, stage, ptrs(inp), ptrs(prev) if prev is not None else None, ptrs(out), out_all.host,
            C.c_void_p(vol.data_ptr()), ft(dt), C.c_void_p(dt_dev.data_ptr() if dt_dev is not None else None),
            C.c_void_p(speed_max.data_ptr() if speed_max is not None else None), C.c_void_p(send_csr[0].data_ptr()),
            C.c_void_p(send_csr[1].data_ptr()), C.c_void_p(send_csr[2].data_ptr()), stream_ptr(stream))
        if rc == 801:
            return False
        check(rc, "fused_stage_push")
        return True

    def stage_part(self, stage, part, inp, prev, out, vol, dt, speed_max=None, stream=None, dt_dev=None):
        """One pass of a stage split in two (t8b200_fused_stage_part): part 1 = the chunks that read no ghost copy,
        part 2 = the partition-boundary chunks.  Returns False when the plan does not support the split."""
        s, ft = _sfx(self.dtype)
        rc = getattr(lib(), "t8b200_fused_stage_part_" + s)(
            self._h, stage, part, ptrs(inp), ptrs(prev) if prev is not None else None, ptrs(out),
            C.c_void_p(vol.data_ptr()), ft(dt), C.c_void_p(dt_dev.data_ptr() if dt_dev is not None else None),
            C.c_void_p(speed_max.data_ptr() if speed_max is not None else None), stream_ptr(stream))
        if rc == 801:
            return False
        check(rc, "fused_stage_part")
        return True

    def stage(self, stage, inp, prev, out, vol, dt, in_all=None, speed_max=None, stream=None, dt_dev=None,
              sync=None):
        """One fused RK stage.  dt_dev: device scalar holding the time step (then `dt` is ignored); sync: a
        PeerMailboxes whose stage epochs order this launch against the peers inside the kernel (boundary chunks wait
        for the previous launch of every peer, the last of them signals this one)."""
        s, ft = _sfx(self.dtype)
        if dt_dev is None and sync is None:
            fn = getattr(lib(), "t8b200_fused_stage_" + s)
            check(fn(self._h, stage, ptrs(inp), in_all.host if in_all is not None else None,
                     ptrs(prev) if prev is not None else None, ptrs(out), C.c_void_p(vol.data_ptr()), ft(dt),
                     C.c_void_p(speed_max.data_ptr() if speed_max is not None else None), stream_ptr(stream)),
                  "fused_stage")
            return
        wait, signal = sync.next_stage() if sync is not None else (0, 0)
        check(getattr(lib(), "t8b200_fused_stage_sync_" + s)(
            self._h, stage, ptrs(inp), in_all.host if in_all is not None else None,
            ptrs(prev) if prev is not None else None, ptrs(out), C.c_void_p(vol.data_ptr()), ft(dt),
            C.c_void_p(dt_dev.data_ptr() if dt_dev is not None else None),
            C.c_void_p(speed_max.data_ptr() if speed_max is not None else None),
            C.byref(sync.struct) if sync is not None else None, C.c_longlong(wait), C.c_longlong(signal),
            stream_ptr(stream)), "fused_stage_sync")


class SubgridPlan:
    """Cell-level tile plan for the fused Subgrid<4,4,4> / Subgrid<4,4> stage kernel."""

    def __init__(self, conn, volumes, dtype, ghost_tail=False):
        import numpy as np
        _torch()
        s, _ = _sfx(dtype)
        npdt = np.float64 if s == "f64" else np.float32
        self.dtype = dtype
        dim = int(conn["dim"])
        vols = np.ascontiguousarray(volumes, dtype=npdt)

        def arr(k, dt):
            v = conn.get(k)
            return None if v is None or len(v) == 0 else np.ascontiguousarray(v, dtype=dt)

        def p(a):
            return None if a is None else a.ctypes.data_as(C.c_void_p)

        keep = [arr("face_neighbors", np.int32), arr("face_normals", npdt), arr("face_areas", npdt),
                arr("level_diff", np.int32), arr("offsets", np.int32), arr("ranks", np.int32),
                arr("indices", np.int32), arr("x_face_neighbors", np.int32), arr("x_face_normals", npdt),
                arr("x_face_areas", npdt), arr("x_level_diff", np.int32), arr("x_offsets", np.int32)]
        h = C.c_void_p()
        create = lib().t8b200_subgrid_plan_create_ghost_tail if ghost_tail else lib().t8b200_subgrid_plan_create
        check(create(
            C.byref(h), int(s == "f64"), dim, C.c_int64(int(conn["n_local"])), C.c_int64(int(conn.get("n_ghost", 0))),
            int(conn["n_faces"]), int(conn["n_bfaces"]), p(keep[0]), p(keep[1]), p(keep[2]), p(keep[3]), p(keep[4]),
            p(vols), p(keep[5]), p(keep[6]), int(conn.get("n_xfaces", 0)), p(keep[7]), p(keep[8]), p(keep[9]),
            p(keep[10]), p(keep[11])), "subgrid_plan_create")
        self._h = h
        lib().t8b200_subgrid_plan_base.restype = C.c_void_p
        self._base = C.c_void_p(lib().t8b200_subgrid_plan_base(h))
        self.n_tail = int(lib().t8b200_plan_ghost_tail_count(self._base))
        info = (C.c_int64 * 8)()
        check(lib().t8b200_subgrid_plan_info(self._h, info), "subgrid_plan_info")
        self.info = dict(n_chunks=info[0], max_halo=info[1], max_faces=info[2], smem_bytes=info[3],
                         device_bytes=info[4], face_records=info[5], halo_entries=info[6], chunk=info[7])

    @classmethod
    def from_device(cls, conn, volumes, dtype, ghost_tail=False, stream=None):
        """Cell-level plan built on the device from DEVICE tensors (t8b200_subgrid_plan_create_device: three kernels
        for structured-only forests, else the generic builder over the cell faces); None only for a rank without
        elements (then use SubgridPlan(conn_to_host(conn), ...))."""
        _torch()
        s, _ = _sfx(dtype)

        def p(k):
            t = conn.get(k)
            return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())

        nx, ng = int(conn.get("n_xfaces", 0)), int(conn.get("n_ghost", 0))
        h = C.c_void_p()
        rc = lib().t8b200_subgrid_plan_create_device(
            C.byref(h), int(s == "f64"), int(conn["dim"]), int(bool(ghost_tail)), C.c_int64(int(conn["n_local"])),
            C.c_int64(ng), int(conn["n_faces"]), int(conn["n_bfaces"]), p("face_neighbors"), p("face_normals"),
            p("face_areas"), p("level_diff"), p("offsets"), C.c_void_p(volumes.data_ptr()), p("ranks") if ng else None,
            p("indices") if ng else None, nx, p("x_face_neighbors") if nx else None, p("x_face_normals") if nx else None,
            p("x_face_areas") if nx else None, p("x_level_diff") if nx else None, p("x_offsets") if nx else None,
            stream_ptr(stream))
        if rc == 801:
            return None
        check(rc, "subgrid_plan_create_device")
        self = cls.__new__(cls)
        self.dtype, self._h = dtype, h
        lib().t8b200_subgrid_plan_base.restype = C.c_void_p
        self._base = C.c_void_p(lib().t8b200_subgrid_plan_base(h))
        self.n_tail = int(lib().t8b200_plan_ghost_tail_count(self._base))
        info = (C.c_int64 * 8)()
        check(lib().t8b200_subgrid_plan_info(self._h, info), "subgrid_plan_info")
        self.info = dict(n_chunks=info[0], max_halo=info[1], max_faces=info[2], smem_bytes=info[3],
                         device_bytes=info[4], face_records=info[5], halo_entries=info[6], chunk=info[7],
                         built_on="device")
        return self

    def device_array(self, which):
        import numpy as np
        n = lib().t8b200_plan_device_array(self._base, which, None, C.c_int64(0))
        if n < 0:
            raise CudaError("plan_device_array(%d)" % which)
        a = np.zeros(n, np.int32)
        if n:
            lib().t8b200_plan_device_array(self._base, which, a.ctypes.data_as(C.c_void_p), C.c_int64(n))
        return a

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            lib().t8b200_subgrid_plan_destroy(self._h)
            self._h = None

    def pull(self, rows, rows_all, stream=None):
        """Ghost-cell tail <- the owners' cell rows (t8b200_ghost_pull on the cell-level plan)."""
        s, _ = _sfx(self.dtype)
        check(getattr(lib(), "t8b200_ghost_pull_" + s)(self._base, len(rows), ptrs(rows), rows_all.host,
                                                       stream_ptr(stream)), "ghost_pull")

    def stage(self, stage, inp, prev, out, vol, dt, in_all=None, stream=None, dt_dev=None, sync=None):
        s, ft = _sfx(self.dtype)
        if dt_dev is None and sync is None:
            check(getattr(lib(), "t8b200_subgrid_fused_stage_" + s)(
                self._h, stage, ptrs(inp), in_all.host if in_all is not None else None,
                ptrs(prev) if prev is not None else None, ptrs(out), C.c_void_p(vol.data_ptr()), ft(dt),
                stream_ptr(stream)), "subgrid_fused_stage")
            return
        wait, signal = sync.next_stage() if sync is not None else (0, 0)
        check(getattr(lib(), "t8b200_subgrid_fused_stage_sync_" + s)(
            self._h, stage, ptrs(inp), in_all.host if in_all is not None else None,
            ptrs(prev) if prev is not None else None, ptrs(out), C.c_void_p(vol.data_ptr()), ft(dt),
            C.c_void_p(dt_dev.data_ptr() if dt_dev is not None else None),
            C.byref(sync.struct) if sync is not None else None, C.c_longlong(wait), C.c_longlong(signal),
            stream_ptr(stream)), "subgrid_fused_stage_sync")


# ------------------------------------------------------------------------------------------------ Cartesian meshes

class _RawDeviceArray:
    """Exposes a library-owned device buffer through __cuda_array_interface__ so torch can copy it."""
    _TYPESTR = {"torch.int32": "<i4", "torch.float32": "<f4", "torch.float64": "<f8", "torch.int64": "<i8"}

    def __init__(self, ptr, n, dtype):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr=self._TYPESTR[str(dtype)],
                                             data=(int(ptr), False), version=2, strides=None)


class _CartConn(C.Structure):
    _fields_ = [("n_local", C.c_int64), ("n_ghost", C.c_int64), ("n_faces", C.c_int64), ("n_bfaces", C.c_int64),
                ("n_xfaces", C.c_int64), ("ranks", C.c_void_p), ("indices", C.c_void_p),
                ("face_neighbors", C.c_void_p), ("face_normals", C.c_void_p), ("face_surfaces", C.c_void_p),
                ("volumes", C.c_void_p), ("centroids", C.c_void_p), ("x_face_neighbors", C.c_void_p),
                ("x_face_normals", C.c_void_p), ("x_face_surfaces", C.c_void_p)]


def cartesian_uniform_connectivity(dim, level, dtype, nranks=1, rank=0, device=None, brick=(1, 1, 1)):
    """Device-built connectivity of a uniform periodic quad/hex forest (a brick of unit trees, default one tree) in
    the reference layout -> dict of torch tensors (copied out of the library-owned buffers)."""
    torch = _torch()
    s, _ = _sfx(dtype)
    device = device or torch.device("cuda", torch.cuda.current_device())
    c = _CartConn()
    check(lib().t8b200_cartesian_brick_connectivity(C.byref(c), int(s == "f64"), dim, level, int(brick[0]),
                                                    int(brick[1]), int(brick[2]), nranks, rank, stream_ptr()),
          "cartesian_brick_connectivity")

    def grab(ptr, n, dt):
        if n == 0:
            return torch.empty(0, dtype=dt, device=device)
        return torch.as_tensor(_RawDeviceArray(ptr, n, dt), device=device).clone()

    nl, ng, nf, nx = c.n_local, c.n_ghost, c.n_faces, c.n_xfaces
    out = dict(dim=dim, level=level, n_local=nl, n_ghost=ng, n_faces=nf, n_bfaces=0, n_xfaces=nx, rank=rank,
               nranks=nranks,
               ranks=grab(c.ranks, nl + ng, torch.int32), indices=grab(c.indices, nl + ng, torch.int32),
               face_neighbors=grab(c.face_neighbors, 2 * nf, torch.int32),
               face_normals=grab(c.face_normals, 3 * nf, dtype), face_areas=grab(c.face_surfaces, nf, dtype),
               volumes=grab(c.volumes, nl, dtype), centroids=grab(c.centroids, 3 * nl, dtype),
               x_face_neighbors=grab(c.x_face_neighbors, 2 * nx, torch.int32),
               x_face_normals=grab(c.x_face_normals, 3 * nx, dtype), x_face_areas=grab(c.x_face_surfaces, nx, dtype))
    torch.cuda.synchronize()
    lib().t8b200_cartesian_connectivity_free(C.byref(c))
    return out


def morton_keys(dim, levels, centroids):
    """Morton keys (numpy uint64) of the leaves' anchors at resolution 2^-20 per axis, from their levels and centroids --
    the input of forest_connectivity (with t8code: the anchor coordinates of the elements)."""
    import numpy as np
    lv = np.asarray(levels, np.int64)
    h = np.ldexp(1.0, -lv)
    key = np.zeros(len(lv), np.uint64)
    for d in range(dim):
        c = np.rint((np.asarray(centroids)[:, d] - 0.5 * h) * (1 << 20)).astype(np.uint64)
        for b in range(20):
            key |= ((c >> np.uint64(b)) & np.uint64(1)) << np.uint64(dim * b + d)
    return key


class _SubgridFaceInfo(C.Structure):
    _fields_ = [("level_diff", C.c_void_p), ("offsets", C.c_void_p), ("x_level_diff", C.c_void_p),
                ("x_offsets", C.c_void_p)]


def forest_connectivity(dim, periodic, keys, levels, dtype, nranks=1, rank=0, device=None, subgrid=False):
    """Device-built connectivity of an adaptive 2:1-balanced one-tree forest (t8b200_forest_connectivity, or
    t8b200_forest_subgrid_connectivity with subgrid=True: `dim` normal components, level_diff / offsets per face) in the
    reference layout -> dict of torch tensors.  keys: uint64 / int64 array or tensor, levels: int32, over ALL leaves."""
    import numpy as np
    torch = _torch()
    s, _ = _sfx(dtype)
    device = device or torch.device("cuda", torch.cuda.current_device())
    k = keys if isinstance(keys, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(keys).view(np.int64))
    lv = levels if isinstance(levels, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(levels, dtype=np.int32))
    k, lv = k.to(device), lv.to(torch.int32).to(device)
    c, info = _CartConn(), _SubgridFaceInfo()
    args = (int(s == "f64"), dim, int(bool(periodic)), C.c_int64(k.numel()), C.c_void_p(k.data_ptr()),
            C.c_void_p(lv.data_ptr()), nranks, rank, stream_ptr())
    if subgrid:
        check(lib().t8b200_forest_subgrid_connectivity(C.byref(c), C.byref(info), *args), "forest_subgrid_connectivity")
    else:
        check(lib().t8b200_forest_connectivity(C.byref(c), *args), "forest_connectivity")

    def grab(ptr, n, dt):
        if n == 0:
            return torch.empty(0, dtype=dt, device=device)
        return torch.as_tensor(_RawDeviceArray(ptr, n, dt), device=device).clone()

    nl, ng, nf, nb, nx = c.n_local, c.n_ghost, c.n_faces, c.n_bfaces, c.n_xfaces
    nd = dim if subgrid else 3
    out = dict(dim=dim, n_local=nl, n_ghost=ng, n_faces=nf, n_bfaces=nb, n_xfaces=nx, rank=rank, nranks=nranks,
               ranks=grab(c.ranks, nl + ng, torch.int32), indices=grab(c.indices, nl + ng, torch.int32),
               face_neighbors=grab(c.face_neighbors, 2 * nf + nb, torch.int32),
               face_normals=grab(c.face_normals, nd * (nf + nb), dtype), face_areas=grab(c.face_surfaces, nf + nb, dtype),
               volumes=grab(c.volumes, nl, dtype), centroids=grab(c.centroids, 3 * nl, dtype),
               x_face_neighbors=grab(c.x_face_neighbors, 2 * nx, torch.int32),
               x_face_normals=grab(c.x_face_normals, nd * nx, dtype), x_face_areas=grab(c.x_face_surfaces, nx, dtype))
    if subgrid:
        out.update(level_diff=grab(info.level_diff, nf, torch.int32), offsets=grab(info.offsets, dim * nf, torch.int32),
                   x_level_diff=grab(info.x_level_diff, nx, torch.int32),
                   x_offsets=grab(info.x_offsets, dim * nx, torch.int32))
    torch.cuda.synchronize()
    lib().t8b200_cartesian_connectivity_free(C.byref(c))
    lib().t8b200_subgrid_face_info_free(C.byref(info))
    return out


def adapt_remap(adapt_data, vars_old, vars_new, vol_old, vol_new, subgrid_dim=0, stream=None):
    """adapt_variables_and_volume (subgrid_dim = 0) / subgrid adapt_variables + adapt_volume (3 or 2): device-side remap
    after t8code adapt.  adapt_data: device int32 tensor (n_new + 1); vars_*: lists of per-variable device tensors."""
    s, _ = _sfx(vol_old.dtype)
    check(getattr(lib(), "t8b200_adapt_remap_" + s)(
        int(subgrid_dim), len(vars_old), C.c_int64(adapt_data.numel() - 1), C.c_void_p(adapt_data.data_ptr()),
        ptrs(vars_old), ptrs(vars_new), C.c_void_p(vol_old.data_ptr()), C.c_void_p(vol_new.data_ptr()),
        stream_ptr(stream)), "adapt_remap")


def partition_remap(ranks, indices, vars_new, vars_old_all, vol_new, vol_old_all, cells_per_element=1, stream=None):
    """partition_data / partition_variable_data + partition_volume_data.  ranks / indices: device int32 tensors;
    vars_old_all: RankTables / PointerTables; vol_old_all: device int64 tensor of per-rank volume pointers or None."""
    s, _ = _sfx(vars_new[0].dtype)
    check(getattr(lib(), "t8b200_partition_remap_" + s)(
        len(vars_new), C.c_int64(ranks.numel()), int(cells_per_element), C.c_void_p(ranks.data_ptr()),
        C.c_void_p(indices.data_ptr()), ptrs(vars_new), vars_old_all.host,
        C.c_void_p(vol_new.data_ptr() if vol_new is not None else None),
        C.c_void_p(vol_old_all.data_ptr() if vol_old_all is not None else None), stream_ptr(stream)),
        "partition_remap")


def gradient_criteria(plan, rho, vol, rho_all=None, out=None, stream=None):
    """estimate_gradient + compute_refinement_criteria of the unstructured example on the tile plan `plan` (Plan).
    rho, vol: device tensors; rho_all: device int64 tensor of per-rank density pointers (multi-rank)."""
    torch = _torch()
    s, _ = _sfx(rho.dtype)
    if out is None:
        out = torch.empty(rho.numel(), dtype=rho.dtype, device=rho.device)
    check(getattr(lib(), "t8b200_gradient_criteria_" + s)(
        plan._h, C.c_void_p(rho.data_ptr()), C.c_void_p(rho_all.data_ptr() if rho_all is not None else None),
        C.c_void_p(vol.data_ptr()), C.c_void_p(out.data_ptr()), stream_ptr(stream)), "gradient_criteria")
    return out


def subgrid_criteria(dim, rho, vol, out=None, stream=None):
    """compute_refinement_criteria<Subgrid>: H1 seminorm of the density per element / volume."""
    torch = _torch()
    s, _ = _sfx(rho.dtype)
    if out is None:
        out = torch.empty(vol.numel(), dtype=rho.dtype, device=rho.device)
    check(getattr(lib(), "t8b200_subgrid_criteria_" + s)(
        int(dim), C.c_int64(vol.numel()), C.c_void_p(rho.data_ptr()), C.c_void_p(vol.data_ptr()),
        C.c_void_p(out.data_ptr()), stream_ptr(stream)), "subgrid_criteria")
    return out


def ghost_push(src_idx, dst_rank, dst_idx, rows, rows_all, stream=None):
    """Owner-side push of the ghost copies (t8b200_ghost_push): rows_all[k][dst_rank[e]][dst_idx[e]] = rows[k][src_idx[e]]."""
    s, _ = _sfx(rows[0].dtype)
    check(getattr(lib(), "t8b200_ghost_push_" + s)(len(rows), C.c_int64(src_idx.numel()), C.c_void_p(src_idx.data_ptr()),
                                                   C.c_void_p(dst_rank.data_ptr()), C.c_void_p(dst_idx.data_ptr()),
                                                   ptrs(rows), rows_all.host, stream_ptr(stream)), "ghost_push")


def subgrid_z_order(dim, cells, out=None, stream=None):
    """column_major_to_z_order + widening to double (output path of SubgridMeshManager::save_variable_to_vtk): cells =
    device tensor of n_elements * 64 (16) values of one variable -> float64 tensor in the Morton order of the cells."""
    torch = _torch()
    s, _ = _sfx(cells.dtype)
    S = 64 if dim == 3 else 16
    if out is None:
        out = torch.empty(cells.numel(), dtype=torch.float64, device=cells.device)
    check(getattr(lib(), "t8b200_subgrid_z_order_" + s)(int(dim), C.c_int64(cells.numel() // S),
                                                        C.c_void_p(cells.data_ptr()), C.c_void_p(out.data_ptr()),
                                                        stream_ptr(stream)), "subgrid_z_order")
    return out


class SharedBuffer:
    """A device allocation that other processes on this node can map (cudaIpc handle exchange is up to the caller).
    `.tensor(shape, dtype)` views it as a torch tensor; `.handle` is the 64-byte IPC handle."""

    def __init__(self, nbytes, device):
        torch = _torch()
        self.device = device
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        h = (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            check(lib().t8b200_shared_alloc(C.c_size_t(self.nbytes), C.byref(p), h), "shared_alloc")
        self.ptr = p.value
        self.handle = bytes(h)
        self._peers = []

    def tensor(self, shape, dtype):
        torch = _torch()
        n = 1
        for d in shape:
            n *= int(d)
        t = torch.as_tensor(_RawDeviceArray(self.ptr, n, dtype), device=self.device)
        return t.view(*shape)

    def open_peer(self, handle):
        """Map another process's buffer; returns the peer device pointer (int)."""
        torch = _torch()
        p = C.c_void_p()
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        with torch.cuda.device(self.device):
            check(lib().t8b200_shared_open(h, C.byref(p)), "shared_open")
        self._peers.append(p.value)
        return p.value

    def close(self):
        for p in self._peers:
            lib().t8b200_shared_close(C.c_void_p(p))
        self._peers = []
        if self.ptr:
            lib().t8b200_shared_free(C.c_void_p(self.ptr))
            self.ptr = None


class _StageSyncStruct(C.Structure):
    _fields_ = [("nranks", C.c_int), ("rank", C.c_int), ("mailboxes_dev", C.c_void_p), ("counter_dev", C.c_void_p)]


class PeerMailboxes:
    """Mailboxes of t8b200_peer_barrier / t8b200_fused_stage_sync_*: one SharedBuffer of 4 * nranks 16-byte slots per
    rank, mapped on every rank (layout in csrc/peer_sync.cuh).
    `exchange(handles)`: handles = list over ranks of the 64-byte IPC handles (this rank's own entry is ignored).
    Two epoch sequences, consecutive per class on every rank: `stage_epoch` (barriers without a value and the stage
    kernels that order themselves) and `cfl_epoch` (barriers that carry the max wave speed)."""

    def __init__(self, rank, nranks, device):
        torch = _torch()
        self.rank, self.nranks, self.device = rank, nranks, device
        # a whole 2 MiB granule of its own, so that the IPC handle maps exactly this buffer
        self.buf = SharedBuffer(max(64 * nranks, 2 << 20), device)
        self.stage_epoch = 0
        self.cfl_epoch = 0
        self.table = None
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.push_counter = torch.zeros(1, dtype=torch.int32, device=device)   # CTAs of a push + barrier launch
        self.struct = None

    @property
    def handle(self):
        return self.buf.handle

    def exchange(self, handles):
        torch = _torch()
        ptrs_ = [self.buf.ptr if r == self.rank else self.buf.open_peer(handles[r]) for r in range(self.nranks)]
        self.set_table(ptrs_)

    def set_table(self, ptrs_):
        """ptrs_: mailbox address of every rank as seen from this rank's device."""
        torch = _torch()
        self.table = torch.tensor(ptrs_, dtype=torch.int64).to(self.device)
        self.struct = _StageSyncStruct(self.nranks, self.rank, self.table.data_ptr(), self.counter.data_ptr())

    def next_stage(self):
        """(wait_epoch, signal_epoch) of the next self-ordering stage launch."""
        self.stage_epoch += 1
        return self.stage_epoch - 1, self.stage_epoch

    def barrier(self, value=None, out=None, stream=None):
        """Barrier kernel on `stream`; with value / out (device scalars) also the maximum over the ranks."""
        is64 = 1
        if value is not None:
            s, _ = _sfx(value.dtype)
            is64 = int(s == "f64")
            self.cfl_epoch += 1
            epoch = self.cfl_epoch
        else:
            self.stage_epoch += 1
            epoch = self.stage_epoch
        check(lib().t8b200_peer_barrier(self.nranks, self.rank, C.c_longlong(epoch),
                                        C.c_void_p(self.table.data_ptr()),
                                        C.c_void_p(value.data_ptr() if value is not None else None), is64,
                                        C.c_void_p(out.data_ptr() if out is not None else None), stream_ptr(stream)),
              "peer_barrier")

    def push_barrier(self, src_idx, dst_rank, dst_idx, rows, rows_all, value=None, out=None, stream=None):
        """ghost_push + barrier in one launch (t8b200_ghost_push_barrier_*); epochs as in `barrier`."""
        s, _ = _sfx(rows[0].dtype)
        if value is not None:
            self.cfl_epoch += 1
            epoch = self.cfl_epoch
        else:
            self.stage_epoch += 1
            epoch = self.stage_epoch
        check(getattr(lib(), "t8b200_ghost_push_barrier_" + s)(
            len(rows), C.c_int64(src_idx.numel()), C.c_void_p(src_idx.data_ptr()), C.c_void_p(dst_rank.data_ptr()),
            C.c_void_p(dst_idx.data_ptr()), ptrs(rows), rows_all.host, C.c_void_p(self.push_counter.data_ptr()),
            self.nranks, self.rank, C.c_longlong(epoch), C.c_void_p(self.table.data_ptr()),
            C.c_void_p(value.data_ptr() if value is not None else None),
            C.c_void_p(out.data_ptr() if out is not None else None), stream_ptr(stream)), "ghost_push_barrier")

    def close(self):
        self.buf.close()


def timestep(speed_max, cfl, length, dt_cap, dt_dev, stream=None):
    """compute_timestep's formula on the device: dt_dev <- min(dt_cap, cfl * length / speed_max) (dt_cap <= 0: none)."""
    s, ft = _sfx(speed_max.dtype)
    check(getattr(lib(), "t8b200_timestep_" + s)(C.c_void_p(speed_max.data_ptr()), ft(cfl), ft(length), ft(dt_cap),
                                                 C.c_void_p(dt_dev.data_ptr()), stream_ptr(stream)), "timestep")


class PointerTables:
    """`[var][rank] -> pointer` device tables built from raw integer pointers (peer-mapped or own)."""

    def __init__(self, ptrs_by_var_rank, device):
        torch = _torch()
        host = torch.tensor(ptrs_by_var_rank, dtype=torch.int64)
        assert host.shape[0] == NVAR
        self.table = host.to(device)
        self.host = (C.c_void_p * NVAR)()
        for k in range(NVAR):
            self.host[k] = self.table[k].data_ptr()


def init_kelvin_helmholtz(dim, centers, u, stream=None):
    """Cartesian Kelvin-Helmholtz state at `centers` (device tensor, 3 per point) into the 5 tensors `u`."""
    s, _ = _sfx(centers.dtype)
    n = centers.numel() // 3
    check(getattr(lib(), "t8b200_init_kelvin_helmholtz_" + s)(dim, C.c_int64(n), C.c_void_p(centers.data_ptr()),
                                                               ptrs(u), stream_ptr(stream)), "init_kh")


def init_spherical_kelvin_helmholtz(centers, u, stream=None):
    """Kelvin-Helmholtz state on the globe (compressible_euler/solver.cu:17-72) at `centers` (device tensor, 3 per
    point) into the 5 tensors `u`."""
    s, _ = _sfx(centers.dtype)
    n = centers.numel() // 3
    check(getattr(lib(), "t8b200_init_spherical_kelvin_helmholtz_" + s)(C.c_int64(n), C.c_void_p(centers.data_ptr()),
                                                                         ptrs(u), stream_ptr(stream)), "init_spherical_kh")


def conn_to_host(conn):
    """torch-tensor connectivity dict -> numpy dict (for Plan)."""
    import torch
    return {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in conn.items()}


def conn_to_device(conn, dtype, device):
    """numpy connectivity dict (oracle layout) -> torch tensors on device, float arrays cast to dtype."""
    import numpy as np
    torch = _torch()
    out = {}
    for k, v in conn.items():
        if isinstance(v, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(v))
            if t.dtype in (torch.float32, torch.float64):
                t = t.to(dtype)
            out[k] = t.to(device)
        else:
            out[k] = v
    return out


from .solver import EulerSolver, SubgridEulerSolver  # noqa: E402,F401
