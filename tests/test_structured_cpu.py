"""CPU checks of the structured-chunk path of the tile plan (box_layout.cuh / structured.cu) through the host-only plan:
which chunks the builder hands to the structured kernel, and that the kernel's index arithmetic -- restated here in
Python, independently of the C++ -- finds, for every element of such a chunk and every direction, exactly the face
neighbour the connectivity names (own slot inside the box, halo list entry across its boundary).  Also the layout's
bank rule: the lower-neighbour reads of a half-warp touch 16 different 8-byte banks."""
import ctypes as C

import numpy as np
import pytest

import oracle
from test_plan_host_cpu import _arr, _lib, _p, host_plan, subgrid_cell_connectivity


# ---------------------------------------------------------------------------------------------- layouts, restated

class Morton:
    NSLOT = 576
    MASK = (0x49, 0x92, 0x24)

    @staticmethod
    def coords(t):
        b = [(t >> i) & 1 for i in range(8)]
        return (b[0] | b[3] << 1 | b[6] << 2, b[1] | b[4] << 1 | b[7] << 2, b[2] | b[5] << 1)

    @staticmethod
    def index(x, y, z):
        t = 0
        for i in range(3):
            t |= ((x >> i) & 1) << (3 * i)
            t |= ((y >> i) & 1) << (3 * i + 1)
        for i in range(2):
            t |= ((z >> i) & 1) << (3 * i + 2)
        return t

    @staticmethod
    def plane_index(c, d):   # compact index of a boundary element in the box face orthogonal to d
        x, y, z = c
        if d == 0:   # bits y0 z0 y1 z1 y2
            return (y & 1) | (z & 1) << 1 | ((y >> 1) & 1) << 2 | ((z >> 1) & 1) << 3 | ((y >> 2) & 1) << 4
        if d == 1:   # bits x0 z0 x1 z1 x2
            return (x & 1) | (z & 1) << 1 | ((x >> 1) & 1) << 2 | ((z >> 1) & 1) << 3 | ((x >> 2) & 1) << 4
        return (x & 1) | (y & 1) << 1 | ((x >> 1) & 1) << 2 | ((y >> 1) & 1) << 3 | ((x >> 2) & 1) << 4 | ((y >> 2) & 1) << 5

    @staticmethod
    def halo_slot(d, side, idx):
        if d == 0:
            return 256 + 16 * (idx >> 2) + 8 + (0 if side else 1) + 2 * (idx & 3)
        if d == 1:
            return 256 + 16 * (8 + (idx >> 3)) + ((idx & 1) | (0 if side else 2) | ((idx >> 1) & 3) << 2)
        return 256 + 16 * (12 + (idx >> 3)) + ((idx & 3) | (0 if side else 4) | ((idx >> 2) & 1) << 3)


class Subgrid:
    NSLOT = 640

    @staticmethod
    def coords(t):
        el, c = t >> 6, t & 63
        return (4 * (el & 1) + (c & 3), 4 * (el >> 1) + ((c >> 2) & 3), c >> 4)

    @staticmethod
    def index(x, y, z):
        return ((x >> 2) + 2 * (y >> 2)) * 64 + (x & 3) + 4 * (y & 3) + 16 * z

    @staticmethod
    def plane_index(c, d):
        x, y, z = c
        if d == 0:
            return (y & 3) | z << 2 | (y >> 2) << 4
        if d == 1:
            return (x & 3) | z << 2 | (x >> 2) << 4
        return (x & 3) | (y & 3) << 2 | ((x >> 2) + 2 * (y >> 2)) << 4

    @staticmethod
    def halo_slot(d, side, idx):
        if d == 0:
            return 256 + 16 * (idx >> 2) + (0 if side else 3) + 4 * (idx & 3)
        if d == 1:
            return 256 + 16 * (8 + (idx >> 2)) + (0 if side else 12) + (idx & 3)
        return 256 + 16 * ((20 if side else 16) + (idx >> 4)) + (idx & 15)


EXT = (8, 8, 4)


def halo_slots_in_thread_order(L):
    """thread_slot(h) of box_layout.cuh: the h-th used halo slot in ascending order."""
    used = sorted(L.halo_slot(d, s, i) for d in range(3) for s in range(2) for i in range(64 if d == 2 else 32))
    assert len(used) == 256 and len(set(used)) == 256 and used[0] >= 256 and used[-1] < L.NSLOT
    return used


@pytest.mark.parametrize("L", [Morton, Subgrid])
def test_layout_is_a_bijection_and_bank_conflict_free(L):
    assert sorted(L.index(*L.coords(t)) for t in range(256)) == list(range(256))
    assert all(L.index(*L.coords(t)) == t for t in range(256))
    halo_slots_in_thread_order(L)
    for d in range(3):
        for hw in range(16):   # lanes of one half-warp read their lower neighbour along d: 16 distinct 8-byte banks
            banks = []
            for t in range(16 * hw, 16 * hw + 16):
                c = list(L.coords(t))
                if c[d] == 0:
                    banks.append(L.halo_slot(d, 0, L.plane_index(c, d)) % 16)
                else:
                    c[d] -= 1
                    banks.append(L.index(*c) % 16)
            assert len(set(banks)) == 16, (L.__name__, d, hw)


# ---------------------------------------------------------------------------------------------- plans

def structured_arrays(lib, h):
    out = {}
    for which, name in ((0, "hdr"), (13, "s_rec"), (14, "s_halo"), (15, "s_hrank"), (16, "g_list"), (8, "area_tab"),
                        (4, "face_ai"), (1, "halo_elem"), (2, "halo_rank")):
        data, count, eb = C.c_void_p(), C.c_int64(), C.c_int()
        assert lib.t8b200_plan_host_array(h, which, C.byref(data), C.byref(count), C.byref(eb)) == 0
        n = count.value
        dt = {1: np.uint8, 4: np.int32, 8: np.float64}[eb.value]
        out[name] = (np.frombuffer((C.c_char * (n * eb.value)).from_address(data.value), dtype=dt).copy() if n
                     else np.zeros(0, dt))
    return out


def neighbour_table(conn):
    """(element, direction d, side) -> list of (neighbour id or -1 for a wall, area) from the reference-layout arrays."""
    nl, nf, nb = int(conn["n_local"]), int(conn["n_faces"]), int(conn["n_bfaces"])
    nbr = np.asarray(conn["face_neighbors"], np.int64)
    nrm = np.asarray(conn["face_normals"], np.float64).reshape(nf + nb, -1)
    area = np.asarray(conn["face_areas"], np.float64)
    tab = {}

    def add(l, r, n, a):
        d = int(np.argmax(np.abs(n)))
        up = n[d] > 0
        if l < nl:
            tab.setdefault((l, d, 1 if up else 0), []).append((r, a))
        if 0 <= r < nl:
            tab.setdefault((r, d, 0 if up else 1), []).append((l, a))

    for f in range(nf):
        add(int(nbr[2 * f]), int(nbr[2 * f + 1]), nrm[f], area[f])
    for f in range(nb):
        add(int(nbr[2 * nf + f]), -1, nrm[nf + f], area[nf + f])
    nx = int(conn.get("n_xfaces", 0))
    if nx:
        xn = np.asarray(conn["x_face_neighbors"], np.int64)
        xr = np.asarray(conn["x_face_normals"], np.float64).reshape(nx, -1)
        xa = np.asarray(conn["x_face_areas"], np.float64)
        for f in range(nx):
            add(int(xn[2 * f]), int(xn[2 * f + 1]), xr[f], xa[f])
    return tab


def check_structured(conn, A, L, multi=False, expect_some=True, expect_all=False):
    nl = int(conn["n_local"])
    hdr = A["hdr"].reshape(-1, 8)
    nch = len(hdr)
    rec = A["s_rec"].reshape(-1, 4)
    tab = neighbour_table(conn)
    slots = halo_slots_in_thread_order(L)
    thread_of = {s: h for h, s in enumerate(slots)}
    if multi:
        rk, ix = np.asarray(conn["ranks"]), np.asarray(conn["indices"])
        my = int(rk[0])
    s_chunks = set(int(c) for c in rec[:, 2])
    # generic chunk list: chunk id, bit 30 = partition-boundary chunk (multi-rank plans list every generic chunk)
    g_ids = [int(c) & 0x3FFFFFFF for c in A["g_list"]]
    g_bnd = {int(c) & 0x3FFFFFFF: bool(int(c) >> 30) for c in A["g_list"]}
    assert len(set(g_ids)) == len(g_ids)
    assert sorted(s_chunks | set(g_ids)) == (list(range(nch)) if (len(rec) or multi) else sorted(g_ids))
    assert not (s_chunks & set(g_ids))
    if multi:
        # boundary flags: a chunk is flagged iff one of its halo elements lives on another rank; flagged chunks are
        # spread over the first half of their launch (all done and signalled about half way through the kernel)
        hs = len(A["halo_elem"]) // nch
        he, hr = A["halo_elem"].reshape(nch, hs), A["halo_rank"].reshape(nch, hs)
        for c in range(nch):
            want = bool(((he[c] >= 0) & (hr[c] != my)).any())
            got = bool(rec[np.nonzero(rec[:, 2] == c)[0][0], 3]) if c in s_chunks else g_bnd[c]
            assert got == want, (c, got, want)
        all_structured = len(rec) == nch and all(int(hdr[c, 1]) == 256 and int(hdr[c, 0]) == 256 * c for c in range(nch))
        if all_structured:      # such plans keep the element order (chunk b = elements [256 b, 256 b + 256))
            assert [int(c) for c in rec[:, 2]] == list(range(nch))
        for flags in ([bool(x) for x in rec[:, 3]], [g_bnd[c] for c in g_ids]):
            nb, n = sum(flags), len(flags)
            if nb and not all_structured:
                last = max(i for i, f in enumerate(flags) if f)
                assert last < max(nb, n // 2) + 1, (last, nb, n)
    else:
        assert not any(rec[:, 3])
    if expect_all:
        assert len(rec) == nch
    if expect_some:
        assert len(rec) > 0
    # brute-force classification of every chunk
    for c in range(nch):
        e0, cnt = int(hdr[c, 0]), int(hdr[c, 1])
        ok = cnt == 256 and e0 % 256 == 0
        halo, areas = {}, set()
        for t in range(256 if ok else 0):
            co = L.coords(t)
            for d in range(3):
                for side in (0, 1):
                    fs = tab.get((e0 + t, d, side), [])
                    if len(fs) != 1 or fs[0][0] < 0:
                        ok = False
                        continue
                    n, a = fs[0]
                    areas.add(a)
                    inside = co[d] > 0 if side == 0 else co[d] < EXT[d] - 1
                    if inside:
                        cc = list(co)
                        cc[d] += -1 if side == 0 else 1
                        ok = ok and n == e0 + L.index(*cc)
                    else:
                        ok = ok and not (e0 <= n < e0 + 256)
                        halo[L.halo_slot(d, side, L.plane_index(co, d))] = n
        ok = ok and len(areas) == 1 and len(set(halo.values())) == 256
        assert ok == (c in s_chunks), (c, ok)
        if not ok:
            continue
        q = int(np.nonzero(rec[:, 2] == c)[0][0])
        assert rec[q, 0] == e0 and A["area_tab"][rec[q, 1]] == areas.pop()
        lst = A["s_halo"][256 * q:256 * q + 256]
        for slot, n in halo.items():
            h = thread_of[slot]
            if multi:
                assert (int(A["s_hrank"][256 * q + h]), int(lst[h])) == (int(rk[n]), int(ix[n])), (c, slot)
                assert (n >= nl) == (int(rk[n]) != my)
            else:
                assert int(lst[h]) == n, (c, slot)


def _hex(level, periodic=True, amr=False):
    f = oracle.Forest(3, level, periodic=periodic)
    if amr:
        lv, cent, vol, _ = f.elements()
        f = f.adapt(np.where(np.abs(cent[:, 2] - 0.3) < 0.1, 20.0, 0.0), 10.0, 1, level + 1)
    return f


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_uniform_hex_is_all_structured(dtype):
    lib = _lib()
    conn = _hex(4).connectivity(dtype=dtype)
    h, keep = host_plan(lib, conn, dtype)
    try:
        check_structured(conn, structured_arrays(lib, h), Morton, expect_all=True)
    finally:
        lib.t8b200_plan_destroy(h)


@pytest.mark.parametrize("kind", ["amr", "walls", "tiny_periodic", "quad"])
def test_mixed_meshes(kind):
    """Hanging faces, walls and a periodic box as narrow as the chunk leave their chunks to the generic kernel."""
    lib = _lib()
    if kind == "quad":
        f = oracle.Forest(2, 5)
    else:
        f = _hex({"tiny_periodic": 3, "walls": 5}.get(kind, 4), periodic=kind != "walls", amr=kind == "amr")
    conn = f.connectivity(dtype=np.float64)
    h, keep = host_plan(lib, conn, np.float64)
    try:
        A = structured_arrays(lib, h)
        check_structured(conn, A, Morton, expect_some=kind == "walls")
        if kind in ("tiny_periodic", "quad"):
            assert len(A["s_rec"]) == 0
        elif kind == "walls":
            assert 0 < len(A["s_rec"]) // 4 < len(A["hdr"]) // 8
    finally:
        lib.t8b200_plan_destroy(h)


def test_structured_off_switch(monkeypatch):
    lib = _lib()
    conn = _hex(4).connectivity(dtype=np.float64)
    monkeypatch.setenv("T8B200_STRUCTURED", "0")
    # the switch is read once per process: only check that a fresh process honours it
    import subprocess
    import sys
    code = ("import numpy as np, ctypes as C, sys; sys.path.insert(0, 'tests'); import oracle;"
            "from test_plan_host_cpu import _lib, host_plan; from test_structured_cpu import structured_arrays;"
            "lib = _lib(); conn = oracle.Forest(3, 4).connectivity(dtype=np.float64);"
            "h, k = host_plan(lib, conn, np.float64); A = structured_arrays(lib, h);"
            "assert len(A['s_rec']) == 0 and len(A['g_list']) == 0; print('ok')")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True,
                         env=dict(os.environ, T8B200_STRUCTURED="0", PYTHONPATH=root))
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


@pytest.mark.parametrize("P", [2, 3])
def test_partitioned_hex(P):
    lib = _lib()
    f = _hex(4)
    for rank in range(P):
        conn = f.connectivity(P, rank, dtype=np.float64)
        h, keep = host_plan(lib, conn, np.float64)
        try:
            # (a rank whose first element is not box-aligned in the global Morton order has no aligned chunks)
            check_structured(conn, structured_arrays(lib, h), Morton, multi=True, expect_some=P == 2 or rank == 0)
        finally:
            lib.t8b200_plan_destroy(h)


@pytest.mark.parametrize("amr", [False, True])
def test_subgrid_cells(amr):
    """Cell-level plan of Subgrid<4,4,4>: 4 sibling elements = one 8 x 8 x 4 box of cells."""
    lib = _lib()
    f = oracle.Forest(3, 2)
    if amr:
        lv, cent, vol, _ = f.elements()
        f = f.adapt(np.where(cent[:, 2] < 0.25, 1.0, 0.0), 0.02, 1, 3)
    lv, cent, vol, _ = f.elements()
    conn = f.connectivity(subgrid=True, dtype=np.float64)
    keep = [_arr(conn, k, d) for k, d in (("face_neighbors", np.int32), ("face_normals", np.float64),
                                          ("face_areas", np.float64), ("level_diff", np.int32), ("offsets", np.int32))]
    vols = np.ascontiguousarray(vol, np.float64)
    sh = C.c_void_p()
    assert lib.t8b200_subgrid_plan_create_host(C.byref(sh), 1, 3, C.c_int64(f.num_elements), C.c_int64(0),
                                               int(conn["n_faces"]), int(conn["n_bfaces"]), _p(keep[0]), _p(keep[1]),
                                               _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(vols), None, None, 0, None,
                                               None, None, None, None) == 0
    try:
        A = structured_arrays(lib, C.c_void_p(lib.t8b200_subgrid_plan_base(sh)))
        cells = subgrid_cell_connectivity(conn, vol, lv)
        check_structured(cells, A, Subgrid, expect_some=not amr, expect_all=not amr)
    finally:
        lib.t8b200_subgrid_plan_destroy(sh)
