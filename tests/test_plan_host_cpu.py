"""CPU checks of the tile-plan builder (the host logic of the fused path) through the host-only entry points
t8b200_plan_create_host / t8b200_plan_host_array: no device needed.

Every array the stage kernel reads is checked against the connectivity it was built from: a chunk holds exactly the faces
that touch its elements, each once; the halo is the sorted set of outside endpoints; records are in kernel order
(x, y, z, walls; by left slot) with the canonical orientation, the axis in the record and the area through the table;
the element -> face table lists, per element, its records in ascending order with the right side flags."""
import ctypes as C
import os
from collections import Counter

import numpy as np
import pytest

import oracle
from util import hybrid_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EC = 256


def _lib():
    from t8gpu_b200 import build
    L = C.CDLL(build.build())
    L.t8b200_subgrid_plan_base.restype = C.c_void_p
    return L


def _p(a):
    return None if a is None or len(a) == 0 else a.ctypes.data_as(C.c_void_p)


def _arr(conn, k, dt):
    v = conn.get(k)
    return None if v is None or len(v) == 0 else np.ascontiguousarray(v, dtype=dt)


def host_plan(L, conn, dtype):
    keep = [_arr(conn, "face_neighbors", np.int32), _arr(conn, "face_normals", dtype), _arr(conn, "face_areas", dtype),
            _arr(conn, "ranks", np.int32), _arr(conn, "indices", np.int32), _arr(conn, "x_face_neighbors", np.int32),
            _arr(conn, "x_face_normals", dtype), _arr(conn, "x_face_areas", dtype)]
    h = C.c_void_p()
    rc = L.t8b200_plan_create_host(C.byref(h), int(dtype == np.float64), C.c_int64(int(conn["n_local"])),
                                   C.c_int64(int(conn.get("n_ghost", 0))), int(conn["n_faces"]), int(conn["n_bfaces"]),
                                   _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]), _p(keep[4]),
                                   int(conn.get("n_xfaces", 0)), _p(keep[5]), _p(keep[6]), _p(keep[7]))
    assert rc == 0, rc
    return h, keep


def arrays(L, h):
    names = ["hdr", "halo_elem", "halo_rank", "face_lr", "face_ai", "ell", "ovf_off", "ovf_ent", "area_tab", "fnx", "fny",
             "fnz", "farea"]
    dts = {1: np.uint8, 2: np.uint16, 4: np.int32, 8: np.float64}
    out = {}
    for which, name in enumerate(names):
        data, count, eb = C.c_void_p(), C.c_int64(), C.c_int()
        assert L.t8b200_plan_host_array(h, which, C.byref(data), C.byref(count), C.byref(eb)) == 0
        n = count.value
        if n == 0:
            out[name] = np.zeros(0, dts[eb.value])
            continue
        buf = (C.c_char * (n * eb.value)).from_address(data.value)
        a = np.frombuffer(buf, dtype=dts[eb.value]).copy()
        out[name] = a.view(np.uint32) if name == "face_lr" else a
    info = (C.c_int64 * 8)()
    assert L.t8b200_plan_info(h, info) == 0
    out["info"] = list(info)
    return out


def check_plan(conn, dtype, A, multi=False):
    """Brute-force comparison of the plan arrays `A` with the reference-layout connectivity `conn`."""
    nl, nf, nb = int(conn["n_local"]), int(conn["n_faces"]), int(conn["n_bfaces"])
    nx = int(conn.get("n_xfaces", 0))
    nbr = np.asarray(conn["face_neighbors"], np.int64)
    l = np.concatenate([nbr[0:2 * nf:2], nbr[2 * nf:2 * nf + nb], np.asarray(conn.get("x_face_neighbors", []), np.int64)[0::2]])
    r = np.concatenate([nbr[1:2 * nf:2], -np.ones(nb, np.int64), np.asarray(conn.get("x_face_neighbors", []), np.int64)[1::2]])
    nd = len(conn["face_normals"]) // max(1, nf + nb)
    nrm = np.concatenate([np.asarray(conn["face_normals"], np.float64).reshape(-1, nd),
                          np.asarray(conn.get("x_face_normals", np.zeros(0)), np.float64).reshape(-1, nd)])
    nrm = np.pad(nrm, ((0, 0), (0, 3 - nd)))
    area = np.concatenate([np.asarray(conn["face_areas"], dtype).astype(np.float64),
                           np.asarray(conn.get("x_face_areas", np.zeros(0)), dtype).astype(np.float64)])
    ntot = nf + nb + nx
    axis_aligned = np.all((np.abs(nrm) == 1).sum(1) == 1) and np.all((nrm != 0).sum(1) == 1)
    nch = A["info"][0]
    assert nch == len(A["hdr"]) // 8 and (nch > 0) == (nl > 0)
    if nch == 0:
        return
    cmp_mode = len(A["area_tab"]) > 0
    assert cmp_mode == (axis_aligned and len(np.unique(area)) <= 256)
    HS, FS = len(A["halo_elem"]) // nch, len(A["face_lr"]) // nch
    hdr = A["hdr"].reshape(nch, 8)
    ell = A["ell"].reshape(-1, 8)
    # element ranges tile [0, n_local) in order
    assert hdr[0, 0] == 0 and np.all(hdr[1:, 0] == hdr[:-1, 0] + hdr[:-1, 1]) and hdr[-1, 0] + hdr[-1, 1] == nl
    assert np.all(hdr[:, 1] <= EC) and np.all(hdr[:, 1] > 0)
    # faces of every element range, brute force
    touches = [[] for _ in range(nch)]
    start = hdr[:, 0]
    for f in range(ntot):
        cs = set()
        for e in (l[f], r[f]):
            if 0 <= e < nl:
                cs.add(int(np.searchsorted(start, e, side="right") - 1))
        assert cs, "face without a local endpoint"
        for c in cs:
            touches[c].append(f)
    ghost_of = {}
    if multi:
        rk, ix = np.asarray(conn["ranks"]), np.asarray(conn["indices"])
        for g in range(nl, len(rk)):
            ghost_of[(int(rk[g]), int(ix[g]))] = g
    for c in range(nch):
        e0, cnt, nh, nfc = hdr[c, 0], hdr[c, 1], hdr[c, 2] & 0xFFFF, hdr[c, 2] >> 16
        x_end, y_end, z_end = hdr[c, 3] & 0xFFFF, hdr[c, 3] >> 16, hdr[c, 4]
        he = A["halo_elem"][c * HS:(c + 1) * HS]
        assert np.all(he[nh:] == -1)
        if multi:
            hr = A["halo_rank"][c * HS:(c + 1) * HS]
            my = int(np.asarray(conn["ranks"])[0])
            halo_ids = [int(he[i]) if hr[i] == my else ghost_of[(int(hr[i]), int(he[i]))] for i in range(nh)]
        else:
            halo_ids = [int(x) for x in he[:nh]]
        want_halo = sorted({int(e) for f in touches[c] for e in (l[f], r[f]) if e >= 0 and not e0 <= e < e0 + cnt})
        assert halo_ids == want_halo, c
        glob = lambda s: e0 + s if s < EC else halo_ids[s - EC]   # noqa: E731
        lr = A["face_lr"][c * FS:(c + 1) * FS]
        assert nfc == len(touches[c]) and np.all(lr[nfc:] == 0)
        got, want = Counter(), Counter()
        prev_key = None
        per_el = [[] for _ in range(cnt)]
        for j in range(nfc):
            sl, ax, sr = int(lr[j] & 0x3FFF), int((lr[j] >> 14) & 3), int(lr[j] >> 16)
            if cmp_mode:
                wall = j >= z_end
                grp = 3 if wall else (0 if j < x_end else 1 if j < y_end else 2)
                a = A["area_tab"][A["face_ai"][c * FS + j]]
                if wall:
                    code = sr & 7
                    assert sr >> 3 == 0x1FFF and ax == 0
                    got[(glob(sl), -1, code >> 1, 1 if code & 1 else -1, a)] += 1
                else:
                    assert ax == grp
                    got[(glob(sl), glob(sr), ax, 1, a)] += 1
                key = (grp, sl, sr)
            else:
                n3 = (A["fnx"][c * FS + j], A["fny"][c * FS + j], A["fnz"][c * FS + j])
                got[(glob(sl), -1 if sr == 0xFFFF else glob(sr), n3, A["farea"][c * FS + j])] += 1
                key = (0, sl, sr)
            assert prev_key is None or prev_key <= key, (c, j)   # kernel order
            prev_key = key
            if sl < EC:
                per_el[sl].append(j << 1)
            if sr < EC:
                per_el[sr].append((j << 1) | 1)
        for f in touches[c]:
            if cmp_mode:
                axf = int(np.argmax(np.abs(nrm[f])))
                sg = 1 if nrm[f, axf] > 0 else -1
                if r[f] < 0:
                    want[(int(l[f]), -1, axf, sg, area[f])] += 1
                else:   # canonical orientation: the normal of the record is +e_axis
                    a_, b_ = (int(l[f]), int(r[f])) if sg > 0 else (int(r[f]), int(l[f]))
                    want[(a_, b_, axf, 1, area[f])] += 1
            else:
                want[(int(l[f]), int(r[f]), tuple(np.asarray(nrm[f], dtype).astype(np.float64)), area[f])] += 1
        assert got == want, c
        # element -> face table (+ overflow): the records of each element, ascending
        oo, oe = hdr[c, 5], hdr[c, 6]
        for i in range(cnt):
            ent = [int(x) for x in ell[e0 + i] if x != 0xFFFF]
            if oo >= 0:
                off = A["ovf_off"][oo:oo + EC + 1]
                ent += [int(x) for x in A["ovf_ent"][oe + off[i]:oe + off[i + 1]]]
            assert ent == per_el[i], (c, i)
            assert np.all(ell[e0 + i][min(len(per_el[i]), 8):] == 0xFFFF)
        if cmp_mode:
            ai = A["face_ai"][c * FS:c * FS + nfc]
            assert hdr[c, 7] == (ai[0] if nfc and np.all(ai == ai[0]) else -1)


def _forest(kind):
    if kind == "hex3":
        return oracle.Forest(3, 3)
    if kind == "hex3_walls":
        return oracle.Forest(3, 3, periodic=False)
    if kind == "quad5":
        return oracle.Forest(2, 5)
    f = oracle.Forest(3, 2) if kind == "hex_amr" else oracle.Forest(2, 4, periodic=False)
    lv, cent, vol, _ = f.elements()
    crit = np.where(np.abs(cent[:, f.dim - 1] - 0.5) < 0.2, 20.0, 0.0)
    return f.adapt(crit, 10.0, 1, 4 if kind == "hex_amr" else 6)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["hex3", "hex3_walls", "quad5", "hex_amr", "quad_amr_walls"])
def test_plan_arrays_match_connectivity(kind, dtype):
    L = _lib()
    conn = _forest(kind).connectivity(dtype=dtype)
    h, keep = host_plan(L, conn, dtype)
    try:
        check_plan(conn, dtype, arrays(L, h))
    finally:
        L.t8b200_plan_destroy(h)


def test_split_chunks_and_general_normals(monkeypatch):
    L = _lib()
    conn, vol, cent = hybrid_mesh(6, periodic=False, dtype=np.float64)
    h, keep = host_plan(L, conn, np.float64)
    A = arrays(L, h)
    L.t8b200_plan_destroy(h)
    assert len(A["area_tab"]) == 0 and len(A["fnx"]) > 0     # general geometry
    check_plan(conn, np.float64, A)
    # forced splitting: more chunks, same invariants
    f = oracle.Forest(3, 3)
    lv, cent, vol, _ = f.elements()
    conn = f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < 0.2, 20.0, 0.0), 10.0, 1, 4).connectivity(dtype=np.float64)
    h, keep = host_plan(L, conn, np.float64)
    n0 = arrays(L, h)["info"][0]
    L.t8b200_plan_destroy(h)
    monkeypatch.setenv("T8B200_TEST_MAX_HALO", "40")
    h, keep = host_plan(L, conn, np.float64)
    A = arrays(L, h)
    L.t8b200_plan_destroy(h)
    assert A["info"][0] > n0 and A["info"][1] <= 40
    check_plan(conn, np.float64, A)


@pytest.mark.parametrize("P", [2, 3])
def test_partitioned_plan(P):
    """Ghost endpoints go through (rank, index); the faces owned by the lower rank are taken from the x-face arrays."""
    L = _lib()
    f = _forest("hex_amr")
    for rank in range(P):
        conn = f.connectivity(P, rank, dtype=np.float64)
        h, keep = host_plan(L, conn, np.float64)
        try:
            check_plan(conn, np.float64, arrays(L, h), multi=True)
        finally:
            L.t8b200_plan_destroy(h)


def subgrid_cell_connectivity(conn, vol, levels, dim=3):
    """Cell-level connectivity of a Subgrid<4,4,4> (dim 3) / Subgrid<4,4> (dim 2) forest in the reference layout, by
    brute force from the element-level arrays of SubgridMeshConnectivityAccessor: the faces between the cells of an
    element (compute_inner_fluxes, kernels.inl:335-662: area (cbrt(vol)/4)^2 resp. sqrt(vol)/4, normal +e_axis) and
    the 16 (4) sub-faces of every element face (compute_outer_fluxes, kernels.inl:717-802: left cell on the face
    plane, right cell = anchor + (i, j) at full or half stride, area / 16 (4))."""
    ne, nf, nb = int(conn["n_local"]), int(conn["n_faces"]), int(conn["n_bfaces"])
    S, TPF, NJ = (64, 16, 4) if dim == 3 else (16, 4, 1)
    nbr = np.asarray(conn["face_neighbors"], np.int64)
    ld = np.asarray(conn["level_diff"])
    pairs, normals, areas, walls, wnormals, wareas = [], [], [], [], [], []
    flat = lambda c: c[0] + 4 * c[1] + 16 * c[2]   # noqa: E731

    def pad3(v):
        return [float(x) for x in v] + [0.0] * (3 - len(v))

    for e in range(ne):
        h = 0.5 ** int(levels[e]) / 4.0
        a_in = h * h if dim == 3 else h
        for ax in range(dim):
            tang = [d for d in range(dim) if d != ax]
            for p in range(3):
                for j in range(NJ):
                    for i in range(4):
                        cl, cr = [0, 0, 0], [0, 0, 0]
                        cl[ax], cr[ax] = p, p + 1
                        cl[tang[0]] = cr[tang[0]] = i
                        if dim == 3:
                            cl[tang[1]] = cr[tang[1]] = j
                        pairs.append((e * S + flat(cl), e * S + flat(cr)))
                        n = [0.0, 0.0, 0.0]
                        n[ax] = 1.0
                        normals.append(n)
                        areas.append(a_in)

    def outer(el, er, n, a, ldf, o, wall, out_pairs, out_n, out_a):
        ax = int(np.argmax(np.abs(n)))
        tang = [d for d in range(dim) if d != ax]
        ds = 2 if wall or ldf == 0 else 1
        for j in range(NJ):
            for i in range(4):
                cl = [0, 0, 0]
                cl[ax] = 3 if n[ax] > 0 else 0
                cl[tang[0]] = i
                if dim == 3:
                    cl[tang[1]] = j
                if wall:
                    walls.append(el * S + flat(cl))
                    wnormals.append(pad3(n))
                    wareas.append(a / TPF)
                    continue
                cr = [int(x) for x in o] + [0] * (3 - dim)
                cr[tang[0]] += ds * i // 2
                if dim == 3:
                    cr[tang[1]] += ds * j // 2
                out_pairs.append((el * S + flat(cl), er * S + flat(cr)))
                out_n.append(pad3(n))
                out_a.append(a / TPF)

    nrm = np.asarray(conn["face_normals"], np.float64).reshape(-1, dim)
    area = np.asarray(conn["face_areas"], np.float64)
    off = np.asarray(conn["offsets"]).reshape(-1, dim)
    for F in range(nf + nb):
        wall = F >= nf
        el = int(nbr[2 * F]) if not wall else int(nbr[2 * nf + (F - nf)])
        er = int(nbr[2 * F + 1]) if not wall else -1
        outer(el, er, nrm[F], area[F], 0 if wall else ld[F], None if wall else off[F], wall, pairs, normals, areas)
    out = dict(n_local=ne * S, n_ghost=0, n_faces=len(pairs), n_bfaces=len(walls),
               face_neighbors=np.concatenate([np.asarray(pairs, np.int32).reshape(-1), np.asarray(walls, np.int32)]),
               face_normals=np.asarray(normals + wnormals, np.float64).reshape(-1),
               face_areas=np.asarray(areas + wareas, np.float64))
    if int(conn.get("n_ghost", 0)) > 0:   # partition: ghost elements -> S ghost cells each, x-faces -> TPF sub-faces
        ng, nx = int(conn["n_ghost"]), int(conn.get("n_xfaces", 0))
        out["n_ghost"] = ng * S
        out["ranks"] = np.repeat(np.asarray(conn["ranks"], np.int32), S)
        out["indices"] = (np.repeat(np.asarray(conn["indices"], np.int64) * S, S) +
                          np.tile(np.arange(S), ne + ng)).astype(np.int32)
        xn = np.asarray(conn.get("x_face_neighbors", []), np.int64)
        xnrm = np.asarray(conn.get("x_face_normals", []), np.float64).reshape(-1, dim)
        xa, xld = np.asarray(conn.get("x_face_areas", []), np.float64), np.asarray(conn.get("x_level_diff", []))
        xoff = np.asarray(conn.get("x_offsets", [])).reshape(-1, dim)
        xp, xnn, xaa = [], [], []
        for F in range(nx):
            outer(int(xn[2 * F]), int(xn[2 * F + 1]), xnrm[F], xa[F], xld[F], xoff[F], False, xp, xnn, xaa)
        out.update(n_xfaces=len(xp), x_face_neighbors=np.asarray(xp, np.int32).reshape(-1),
                   x_face_normals=np.asarray(xnn, np.float64).reshape(-1), x_face_areas=np.asarray(xaa, np.float64))
    return out


@pytest.mark.parametrize("periodic", [True, False])
def test_subgrid_cell_plan_matches_brute_force(periodic):
    L = _lib()
    f = oracle.Forest(3, 1, periodic=periodic)
    lv, cent, vol, _ = f.elements()
    f = f.adapt(np.where(cent[:, 2] < 0.5, 1.0, 0.0), 0.02, 1, 2)
    lv, cent, vol, _ = f.elements()
    conn = f.connectivity(subgrid=True, dtype=np.float64)
    keep = [_arr(conn, k, d) for k, d in (("face_neighbors", np.int32), ("face_normals", np.float64),
                                          ("face_areas", np.float64), ("level_diff", np.int32), ("offsets", np.int32))]
    vols = np.ascontiguousarray(vol, np.float64)
    sh = C.c_void_p()
    assert L.t8b200_subgrid_plan_create_host(C.byref(sh), 1, 3, C.c_int64(f.num_elements), C.c_int64(0),
                                             int(conn["n_faces"]), int(conn["n_bfaces"]), _p(keep[0]), _p(keep[1]),
                                             _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(vols), None, None, 0, None,
                                             None, None, None, None) == 0
    A = arrays(L, C.c_void_p(L.t8b200_subgrid_plan_base(sh)))
    L.t8b200_subgrid_plan_destroy(sh)
    check_plan(subgrid_cell_connectivity(conn, vol, lv), np.float64, A)


def test_subgrid_cell_plan_2d_matches_brute_force():
    """Subgrid<4,4>: 16 cells per element, 4 sub-faces per element face, 2-component normals and offsets."""
    L = _lib()
    f = oracle.Forest(2, 2, periodic=False)
    lv, cent, vol, _ = f.elements()
    f = f.adapt(np.where(cent[:, 1] < 0.5, 1.0, 0.0), 0.02, 1, 3)
    lv, cent, vol, _ = f.elements()
    conn = f.connectivity(subgrid=True, dtype=np.float64)
    keep = [_arr(conn, k, d) for k, d in (("face_neighbors", np.int32), ("face_normals", np.float64),
                                          ("face_areas", np.float64), ("level_diff", np.int32), ("offsets", np.int32))]
    vols = np.ascontiguousarray(vol, np.float64)
    sh = C.c_void_p()
    assert L.t8b200_subgrid_plan_create_host(C.byref(sh), 1, 2, C.c_int64(f.num_elements), C.c_int64(0),
                                             int(conn["n_faces"]), int(conn["n_bfaces"]), _p(keep[0]), _p(keep[1]),
                                             _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(vols), None, None, 0, None,
                                             None, None, None, None) == 0
    A = arrays(L, C.c_void_p(L.t8b200_subgrid_plan_base(sh)))
    L.t8b200_subgrid_plan_destroy(sh)
    check_plan(subgrid_cell_connectivity(conn, vol, lv, dim=2), np.float64, A)


def test_subgrid_cell_plan_partitioned():
    L = _lib()
    f = oracle.Forest(3, 1)
    lv, cent, vol, _ = f.elements()
    f = f.adapt(np.where(cent[:, 2] < 0.5, 1.0, 0.0), 0.02, 1, 2)
    lv, cent, vol, _ = f.elements()
    P = 2
    off = f.partition_offsets(P)
    for rank in range(P):
        conn = f.connectivity(P, rank, subgrid=True, dtype=np.float64)
        names = [("face_neighbors", np.int32), ("face_normals", np.float64), ("face_areas", np.float64),
                 ("level_diff", np.int32), ("offsets", np.int32), ("ranks", np.int32), ("indices", np.int32),
                 ("x_face_neighbors", np.int32), ("x_face_normals", np.float64), ("x_face_areas", np.float64),
                 ("x_level_diff", np.int32), ("x_offsets", np.int32)]
        k = [_arr(conn, n_, d) for n_, d in names]
        vols = np.ascontiguousarray(vol[off[rank]:off[rank + 1]], np.float64)
        sh = C.c_void_p()
        assert L.t8b200_subgrid_plan_create_host(
            C.byref(sh), 1, 3, C.c_int64(int(conn["n_local"])), C.c_int64(int(conn["n_ghost"])), int(conn["n_faces"]),
            int(conn["n_bfaces"]), _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), _p(vols), _p(k[5]), _p(k[6]),
            int(conn["n_xfaces"]), _p(k[7]), _p(k[8]), _p(k[9]), _p(k[10]), _p(k[11])) == 0
        A = arrays(L, C.c_void_p(L.t8b200_subgrid_plan_base(sh)))
        L.t8b200_subgrid_plan_destroy(sh)
        assert int(conn["n_ghost"]) > 0
        cells = subgrid_cell_connectivity(conn, vols, lv[off[rank]:off[rank + 1]])
        check_plan(cells, np.float64, A, multi=True)


def test_subgrid_cell_plan():
    """Cell-level plan of a Subgrid<4,4,4> forest with hanging faces: chunk statistics and conservation of the face
    count (every cell face appears once per chunk it touches; inner faces of an element never leave its chunk)."""
    L = _lib()
    f = oracle.Forest(3, 1)
    lv, cent, vol, _ = f.elements()
    f = f.adapt(np.where(cent[:, 2] < 0.5, 1.0, 0.0), 0.02, 1, 2)
    lv, cent, vol, _ = f.elements()
    conn = f.connectivity(subgrid=True, dtype=np.float64)
    keep = [_arr(conn, k, d) for k, d in (("face_neighbors", np.int32), ("face_normals", np.float64),
                                          ("face_areas", np.float64), ("level_diff", np.int32), ("offsets", np.int32))]
    vols = np.ascontiguousarray(vol, np.float64)
    sh = C.c_void_p()
    rc = L.t8b200_subgrid_plan_create_host(C.byref(sh), 1, 3, C.c_int64(f.num_elements), C.c_int64(0),
                                           int(conn["n_faces"]), int(conn["n_bfaces"]), _p(keep[0]), _p(keep[1]),
                                           _p(keep[2]), _p(keep[3]), _p(keep[4]), _p(vols), None, None, 0, None, None,
                                           None, None, None)
    assert rc == 0
    A = arrays(L, C.c_void_p(L.t8b200_subgrid_plan_base(sh)))
    nch = A["info"][0]
    ncell = f.num_elements * 64
    # (blocks of 4 elements whose 2:1 halo exceeds the kernel's 256 slots are split)
    assert nch >= (ncell + EC - 1) // EC and A["info"][1] <= 256 and len(A["area_tab"]) >= 2   # two levels: two areas
    hdr = A["hdr"].reshape(nch, 8)
    assert hdr[:, 1].sum() == ncell
    # every cell has at least its 6 faces in the table (more where 2:1 faces hang), all entries valid record numbers
    ell = A["ell"].reshape(-1, 8)
    nent = (ell != 0xFFFF).sum(1)
    assert nent.min() >= 6 and np.all(ell[:, :6] != 0xFFFF)
    FS = len(A["face_lr"]) // nch
    for c in range(nch):
        nfc = hdr[c, 2] >> 16
        e = ell[hdr[c, 0]:hdr[c, 0] + hdr[c, 1]]
        assert (e[e != 0xFFFF] >> 1).max() < nfc <= FS
    L.t8b200_subgrid_plan_destroy(sh)


def test_host_plan_is_not_launchable():
    L = _lib()
    conn = _forest("hex3").connectivity(dtype=np.float64)
    h, keep = host_plan(L, conn, np.float64)
    ptr5 = (C.c_void_p * 5)()
    assert L.t8b200_fused_stage_f64(h, 1, ptr5, None, None, ptr5, C.c_void_p(8), C.c_double(0.1), None, None) != 0
    data, count, eb = C.c_void_p(), C.c_int64(), C.c_int()
    assert L.t8b200_plan_host_array(h, 99, C.byref(data), C.byref(count), C.byref(eb)) != 0
    L.t8b200_plan_destroy(h)


# ---- the device builder's per-block program (csrc/plan_block.cuh) run on the host == the host builder, bit for bit ----
def _plan_with(fn, conn, dtype):
    keep = [_arr(conn, "face_neighbors", np.int32), _arr(conn, "face_normals", dtype), _arr(conn, "face_areas", dtype),
            _arr(conn, "ranks", np.int32), _arr(conn, "indices", np.int32), _arr(conn, "x_face_neighbors", np.int32),
            _arr(conn, "x_face_normals", dtype), _arr(conn, "x_face_areas", dtype)]
    h = C.c_void_p()
    rc = fn(C.byref(h), int(dtype == np.float64), C.c_int64(int(conn["n_local"])), C.c_int64(int(conn.get("n_ghost", 0))),
            int(conn["n_faces"]), int(conn["n_bfaces"]), _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]), _p(keep[4]),
            int(conn.get("n_xfaces", 0)), _p(keep[5]), _p(keep[6]), _p(keep[7]))
    assert rc == 0, rc
    return h, keep


def _int_arrays(L, h, which):
    out = {}
    for w in which:
        data, count, eb = C.c_void_p(), C.c_int64(), C.c_int()
        assert L.t8b200_plan_host_array(h, w, C.byref(data), C.byref(count), C.byref(eb)) == 0
        n = count.value
        out[w] = (np.frombuffer((C.c_char * (n * 4)).from_address(data.value), dtype=np.int32).copy() if n
                  else np.zeros(0, np.int32))
    return out


def _same_plan(L, conn, dtype, multi):
    h1, k1 = _plan_with(L.t8b200_plan_create_host, conn, dtype)
    h2, k2 = _plan_with(L.t8b200_plan_create_block_program_host, conn, dtype)
    A, B = arrays(L, h1), arrays(L, h2)
    assert A["info"] == B["info"]
    for k in A:
        if k != "info":
            assert np.array_equal(A[k], B[k]), k
    if not multi:   # chunk lists of the structured / generic kernels (multi: device post-pass, tests/test_device_plan_gpu.py)
        X, Y = _int_arrays(L, h1, (13, 14, 15, 16)), _int_arrays(L, h2, (13, 14, 15, 16))
        for k in X:
            assert np.array_equal(X[k], Y[k]), k
    L.t8b200_plan_destroy(h1)
    L.t8b200_plan_destroy(h2)
    return A


def _twice_adapted():
    f = oracle.Forest(3, 3)
    for width in (0.2, 0.1):
        lv, cent, vol, _ = f.elements()
        f = f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < width, 20.0, 0.0), 10.0, 1, 5)
    return f


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_block_program_equals_host_builder(dtype):
    L = _lib()
    A = _same_plan(L, oracle.Forest(3, 4).connectivity(dtype=dtype), dtype, False)
    assert A["info"][0] == 16                                             # all structured
    _same_plan(L, oracle.Forest(3, 3, periodic=False).connectivity(dtype=dtype), dtype, False)   # walls
    _same_plan(L, oracle.Forest(2, 5).connectivity(dtype=dtype), dtype, False)                   # quads
    f = _twice_adapted()
    A = _same_plan(L, f.connectivity(dtype=dtype), dtype, False)
    assert len(A["ovf_ent"]) > 0                                          # elements with more than 8 faces
    for P in (2, 3):
        for r in range(P):
            _same_plan(L, f.connectivity(P, r, dtype=dtype), dtype, True)
    for per in (True, False):                                             # general normals, shuffled numbering
        A = _same_plan(L, hybrid_mesh(12, per, dtype, seed=3, shuffle=not per)[0], dtype, False)
        assert len(A["fnx"]) > 0


def test_block_program_splits_like_host_builder(monkeypatch):
    monkeypatch.setenv("T8B200_TEST_MAX_HALO", "40")
    L = _lib()
    f = oracle.Forest(3, 4)
    A = _same_plan(L, f.connectivity(dtype=np.float64), np.float64, False)
    assert A["info"][0] > 16
    for r in range(2):
        _same_plan(L, f.connectivity(2, r, dtype=np.float64), np.float64, True)
    f = oracle.Forest(3, 2)
    for _ in range(3):
        lv, cent, vol, _v = f.elements()
        f = f.adapt(np.where(np.abs(cent[:, 0] - 0.5) + np.abs(cent[:, 1] - 0.5) < 0.3, 20.0, 0.0), 10.0, 1, 5)
    _same_plan(L, f.connectivity(dtype=np.float32), np.float32, False)


def _sg_plan_with(fn, L, conn, vol, dtype, dim):
    keep = [_arr(conn, k, d) for k, d in (("face_neighbors", np.int32), ("face_normals", dtype), ("face_areas", dtype),
                                          ("level_diff", np.int32), ("offsets", np.int32), ("ranks", np.int32),
                                          ("indices", np.int32), ("x_face_neighbors", np.int32), ("x_face_normals", dtype),
                                          ("x_face_areas", dtype), ("x_level_diff", np.int32), ("x_offsets", np.int32))]
    vols = np.ascontiguousarray(vol, dtype)
    ng, nx = int(conn.get("n_ghost", 0)), int(conn.get("n_xfaces", 0))
    sh = C.c_void_p()
    assert fn(C.byref(sh), int(dtype == np.float64), dim, C.c_int64(int(conn["n_local"])), C.c_int64(ng),
              int(conn["n_faces"]), int(conn["n_bfaces"]), _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]), _p(keep[4]),
              _p(vols), _p(keep[5]) if ng else None, _p(keep[6]) if ng else None, nx, _p(keep[7]), _p(keep[8]),
              _p(keep[9]), _p(keep[10]), _p(keep[11])) == 0
    return sh, keep, vols


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_block_program_equals_host_builder_on_subgrid_cells(dtype):
    """The cell-face source of Subgrid<4,4,4> / <4,4> through the device builder's block program (host loops)."""
    L = _lib()
    cases = []
    f = oracle.Forest(3, 2)
    cases.append((f, 3))                                                   # uniform: structured SubgridBox chunks
    lv, cent, vol, _ = f.elements()
    cases.append((f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < 0.3, 20.0, 0.0), 10.0, 1, 4), 3))
    g = oracle.Forest(3, 1, periodic=False)
    lv, cent, vol, _ = g.elements()
    cases.append((g.adapt(np.where(cent[:, 2] < 0.5, 1.0, 0.0), 0.02, 1, 2), 3))      # walls
    h = oracle.Forest(2, 2, periodic=False)
    lv, cent, vol, _ = h.elements()
    cases.append((h.adapt(np.where(cent[:, 1] < 0.5, 1.0, 0.0), 0.02, 1, 3), 2))      # Subgrid<4,4>
    for forest, dim in cases:
        vol = forest.elements()[2]
        for P in (1, 2):
            off = forest.partition_offsets(P)
            for r in range(P):
                conn = forest.connectivity(P, r, subgrid=True, dtype=dtype)
                lvol = vol[off[r]:off[r + 1]]
                s1, k1, v1 = _sg_plan_with(L.t8b200_subgrid_plan_create_host, L, conn, lvol, dtype, dim)
                s2, k2, v2 = _sg_plan_with(L.t8b200_subgrid_plan_create_block_program_host, L, conn, lvol, dtype, dim)
                h1, h2 = C.c_void_p(L.t8b200_subgrid_plan_base(s1)), C.c_void_p(L.t8b200_subgrid_plan_base(s2))
                A, B = arrays(L, h1), arrays(L, h2)
                assert A["info"] == B["info"]
                for k in A:
                    if k != "info":
                        assert np.array_equal(A[k], B[k]), (k, dim, P, r)
                if P == 1:
                    X, Y = _int_arrays(L, h1, (13, 14, 15, 16)), _int_arrays(L, h2, (13, 14, 15, 16))
                    for k in X:
                        assert np.array_equal(X[k], Y[k]), k
                L.t8b200_subgrid_plan_destroy(s1)
                L.t8b200_subgrid_plan_destroy(s2)


def test_block_program_edge_cases():
    """More than 256 distinct areas on axis-aligned faces (no compressed geometry), exactly-fitting area table, meshes
    smaller than one block."""
    L = _lib()
    rng = np.random.default_rng(5)
    for k in (300, 200):
        conn = oracle.Forest(3, 4).connectivity(dtype=np.float64)
        conn["face_areas"] = conn["face_areas"] * (1.0 + rng.integers(0, k, conn["face_areas"].size) / 1024.0)
        A = _same_plan(L, conn, np.float64, False)
        assert (len(A["area_tab"]), len(A["fnx"]) > 0) == ((0, True) if k > 256 else (k, False))
    _same_plan(L, oracle.Forest(3, 2).connectivity(dtype=np.float32), np.float32, False)
    _same_plan(L, oracle.Forest(2, 1).connectivity(dtype=np.float64), np.float64, False)
