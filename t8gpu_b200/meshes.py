"""Synthetic meshes for the harness (tests, bench.py): flat-array mixed-element meshes in the reference's
MeshConnectivityAccessor layout, their periodic tiling to large sizes and their partition into per-rank arrays.
t8code's hybrid cmeshes / simplex and prism schemes are not available in this image (BASELINE config 5 "blocked on
t8code" end to end); the kernels are element-type agnostic, so these arrays exercise exactly what a t8code mixed mesh
would hand them.  Pure numpy; no dependency on the oracle."""
import numpy as np


def hybrid_mesh(n=6, periodic=True, dtype=np.float64, seed=0, shuffle=False, with_shift=False):
    """Synthetic conforming mesh of hexahedra, prisms and tetrahedra (BASELINE config 5 at kernel level: t8code's hybrid
    cmeshes are not available here).  n^3 unit cubes scaled to [0,1]^3, z-layers cycling hex / 2 prisms / 6 Kuhn
    tetrahedra; horizontal quads are split along the (0,0)-(1,1) diagonal so that every interface matches vertex for
    vertex.  Returns (conn, volumes, centroids) with conn in the reference's MeshConnectivityAccessor layout
    (general unit normals pointing left -> right, areas; boundary faces last when not periodic)."""
    assert n % 3 == 0
    h = 1.0 / n
    elems = []   # (origin (i,j,k), list of faces as tuples of local integer vertex coordinates, volume factor, centroid)
    corners = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    c = corners

    def tri(a, b, d):
        return (a, b, d)

    for k in range(n):
        for j in range(n):
            for i in range(n):
                o = (i, j, k)
                kind = k % 3
                if kind == 0:     # hexahedron: 4 side quads, bottom and top as 2 triangles each
                    faces = [(c[0], c[3], c[7], c[4]), (c[1], c[2], c[6], c[5]), (c[0], c[1], c[5], c[4]),
                             (c[3], c[2], c[6], c[7]), tri(c[0], c[1], c[2]), tri(c[0], c[2], c[3]),
                             tri(c[4], c[5], c[6]), tri(c[4], c[6], c[7])]
                    elems.append((o, faces, 1.0, (0.5, 0.5, 0.5)))
                elif kind == 1:   # two prisms over the triangles (0,1,2) and (0,2,3)
                    for t, cen in (((0, 1, 2), (2 / 3, 1 / 3, 0.5)), ((0, 2, 3), (1 / 3, 2 / 3, 0.5))):
                        b = [c[t[0]], c[t[1]], c[t[2]]]
                        u = [c[t[0] + 4], c[t[1] + 4], c[t[2] + 4]]
                        faces = [tuple(b), tuple(u), (b[0], b[1], u[1], u[0]), (b[1], b[2], u[2], u[1]),
                                 (b[2], b[0], u[0], u[2])]
                        elems.append((o, faces, 0.5, cen))
                else:             # Kuhn triangulation: one tetrahedron per permutation of the axes
                    import itertools
                    for perm in itertools.permutations(range(3)):
                        v = [(0, 0, 0)]
                        for ax in perm:
                            w = list(v[-1])
                            w[ax] += 1
                            v.append(tuple(w))
                        faces = [tri(v[1], v[2], v[3]), tri(v[0], v[2], v[3]), tri(v[0], v[1], v[3]), tri(v[0], v[1], v[2])]
                        cen = tuple(sum(p[d] for p in v) / 4.0 for d in range(3))
                        elems.append((o, faces, 1.0 / 6.0, cen))
    order = np.arange(len(elems))
    if shuffle:
        np.random.default_rng(seed).shuffle(order)
    elems = [elems[q] for q in order]

    def key(o, face):   # global (periodic) vertex ids of a face, sorted
        ids = []
        for p in face:
            g = [o[d] + p[d] for d in range(3)]
            if periodic:
                g = [x % n for x in g]
            ids.append((g[0] * (n + 1) + g[1]) * (n + 1) + g[2])
        return tuple(sorted(ids))

    def geom(o, face, cen):   # area vector pointing out of the element
        P = np.array([[o[d] + p[d] for d in range(3)] for p in face], float) * h
        if len(face) == 3:
            a = 0.5 * np.cross(P[1] - P[0], P[2] - P[0])
        else:
            a = 0.5 * np.cross(P[2] - P[0], P[3] - P[1])
        centre = (np.array(o, float) + np.array(cen)) * h
        if np.dot(a, P.mean(0) - centre) < 0:
            a = -a
        return a

    def face_centre(o, face):   # in cell units, not wrapped
        return np.array([[o[d] + p[d] for d in range(3)] for p in face], float).mean(0)

    seen = {}
    nbr, nrm, area, bnbr, bnrm, barea, shift = [], [], [], [], [], [], []
    closed = np.zeros((len(elems), 3))
    for e, (o, faces, vf, cen) in enumerate(elems):
        for f in faces:
            a = geom(o, f, cen)
            closed[e] += a
            kf = key(o, f)
            if kf in seen:
                l, al, cl = seen.pop(kf)
                assert np.allclose(al, -a, atol=1e-14), "non-matching interface"
                # periodic image of the right element that touches the left one: right + shift * domain
                shift.append(np.rint((cl - face_centre(o, f)) / n).astype(np.int64))
                nbr += [l, e]
                A = np.linalg.norm(al)
                nrm += list(al / A)
                area.append(A)
            else:
                seen[kf] = (e, a, face_centre(o, f))
    assert np.abs(closed).max() < 1e-14, "element surfaces are not closed"
    for kf, (e, a, _) in sorted(seen.items(), key=lambda t: t[1][0]):
        assert not periodic, "unmatched face in a periodic mesh"
        A = np.linalg.norm(a)
        bnbr.append(e)
        bnrm += list(a / A)
        barea.append(A)
    vol = np.array([vf * h ** 3 for (_, _, vf, _) in elems])
    cent = np.array([[(o[d] + cen[d]) * h for d in range(3)] for (o, _, _, cen) in elems])
    conn = dict(n_local=len(elems), n_ghost=0, n_faces=len(area), n_bfaces=len(barea),
                face_neighbors=np.array(nbr + bnbr, np.int32), face_normals=np.array(nrm + bnrm, dtype),
                face_areas=np.array(area + barea, dtype))
    if with_shift:
        return conn, vol.astype(dtype), cent, np.array(shift, np.int64).reshape(-1, 3)
    return conn, vol.astype(dtype), cent


def tile_periodic_mesh(conn, vol, cent, shift, T):
    """A periodic mesh (hybrid_mesh(..., periodic=True, with_shift=True)) repeated T x T x T times and scaled back to the
    unit cube: (T^3 x elements) with the same local structure -- large mixed-element meshes without a mesh generator.
    Element order: tile after tile (x fastest), the pattern's order inside a tile."""
    n0, nf = int(conn["n_local"]), int(conn["n_faces"])
    assert int(conn["n_bfaces"]) == 0
    nbr = np.asarray(conn["face_neighbors"], np.int64).reshape(-1, 2)
    t = np.arange(T)
    tx, ty, tz = np.meshgrid(t, t, t, indexing="ij")
    tiles = np.stack([tx.ravel(order="F"), ty.ravel(order="F"), tz.ravel(order="F")], 1)        # x fastest
    tid = lambda c: (c[:, 0] % T) + T * ((c[:, 1] % T) + T * (c[:, 2] % T))                    # noqa: E731
    base = tid(tiles)[:, None] * n0
    left = (base + nbr[None, :, 0]).reshape(-1)
    right = np.empty((T ** 3, nf), np.int64)
    for q in range(T ** 3):
        right[q] = tid(tiles[q][None, :] + shift) * n0 + nbr[:, 1]
    out = dict(n_local=n0 * T ** 3, n_ghost=0, n_faces=nf * T ** 3, n_bfaces=0,
               face_neighbors=np.stack([left, right.reshape(-1)], 1).reshape(-1).astype(np.int32),
               face_normals=np.tile(np.asarray(conn["face_normals"]), T ** 3),
               face_areas=np.tile(np.asarray(conn["face_areas"]) / (T * T), T ** 3).astype(conn["face_areas"].dtype))
    big_vol = np.tile(np.asarray(vol) / T ** 3, T ** 3).astype(vol.dtype)
    big_cent = ((np.asarray(cent)[None, :, :] + tiles[:, None, :]) / T).reshape(-1, 3)
    return out, big_vol, big_cent


def partition_flat_mesh(conn, vol, nranks, rank):
    """Contiguous equal split of a flat single-rank mesh (the element order is the space-filling order) into the arrays
    one rank of MeshManager would hold (mesh_manager.inl:332-481): local elements [off[r], off[r+1]), ghosts = face
    neighbours outside (ascending global id, i.e. grouped by owner), faces whose ghost belongs to a HIGHER rank in the
    main list (the lower rank owns them, mesh_manager.inl:397), the others as x-faces; every face is seen from its local
    element (normal pointing local -> ghost).  Stand-in for t8code's partition of a mixed-element cmesh."""
    n, nf, nb = int(conn["n_local"]), int(conn["n_faces"]), int(conn["n_bfaces"])
    off = (np.arange(nranks + 1, dtype=np.int64) * n) // nranks
    lo, hi = off[rank], off[rank + 1]
    nbr = np.asarray(conn["face_neighbors"], np.int64)
    pairs, bn = nbr[:2 * nf].reshape(-1, 2), nbr[2 * nf:]
    nrm = np.asarray(conn["face_normals"]).reshape(-1, 3)
    ar = np.asarray(conn["face_areas"])
    l_in = (pairs[:, 0] >= lo) & (pairs[:, 0] < hi)
    r_in = (pairs[:, 1] >= lo) & (pairs[:, 1] < hi)
    both = l_in & r_in
    cut_l = l_in & ~r_in            # local element on the left
    cut_r = r_in & ~l_in            # local element on the right: seen from it, the face is flipped
    ghosts = np.unique(np.concatenate([pairs[cut_l, 1], pairs[cut_r, 0]]))
    owner = np.searchsorted(off, ghosts, side="right") - 1
    gid = {int(g): hi - lo + i for i, g in enumerate(ghosts)}
    to_local = lambda a: np.array([gid[int(x)] for x in a], np.int64)          # noqa: E731
    # faces with a ghost, from the local element's side
    cl, cn, ca = pairs[cut_l, 0] - lo, nrm[cut_l], ar[cut_l]
    cg = to_local(pairs[cut_l, 1]) if cut_l.any() else np.zeros(0, np.int64)
    dl, dn, da = pairs[cut_r, 1] - lo, -nrm[cut_r], ar[cut_r]
    dg = to_local(pairs[cut_r, 0]) if cut_r.any() else np.zeros(0, np.int64)
    gl = np.concatenate([cl, dl]); gg = np.concatenate([cg, dg])
    gn = np.concatenate([cn, dn]); ga = np.concatenate([ca, da])
    g_owner = owner[gg - (hi - lo)] if len(gg) else np.zeros(0, np.int64)
    mine = g_owner > rank
    order = np.argsort(gl, kind="stable")        # by local element, as the reference's element loop emits them
    gl, gg, gn, ga, mine = gl[order], gg[order], gn[order], ga[order], mine[order]
    bmask = (bn >= lo) & (bn < hi)
    main_nbr = np.concatenate([np.stack([pairs[both, 0] - lo, pairs[both, 1] - lo], 1),
                               np.stack([gl[mine], gg[mine]], 1)]).reshape(-1)
    out = dict(n_local=int(hi - lo), n_ghost=len(ghosts), n_faces=int(both.sum() + mine.sum()), n_bfaces=int(bmask.sum()),
               rank=rank, nranks=nranks,
               ranks=np.concatenate([np.full(hi - lo, rank), owner]).astype(np.int32),
               indices=np.concatenate([np.arange(hi - lo), ghosts - off[owner]]).astype(np.int32),
               face_neighbors=np.concatenate([main_nbr, bn[bmask] - lo]).astype(np.int32),
               face_normals=np.concatenate([nrm[:nf][both], gn[mine], nrm[nf:][bmask]]).reshape(-1).astype(nrm.dtype),
               face_areas=np.concatenate([ar[:nf][both], ga[mine], ar[nf:][bmask]]).astype(ar.dtype),
               n_xfaces=int((~mine).sum()),
               x_face_neighbors=np.stack([gl[~mine], gg[~mine]], 1).reshape(-1).astype(np.int32),
               x_face_normals=gn[~mine].reshape(-1).astype(nrm.dtype), x_face_areas=ga[~mine].astype(ar.dtype),
               offsets_global=off)
    return out, np.asarray(vol)[lo:hi]


def smooth_state(cent, dtype, seed=0, amp=0.05):
    """Smooth density / velocity / pressure field at the given points plus a seeded perturbation."""
    rng = np.random.default_rng(seed)
    n = len(cent)
    x, y, z = cent[:, 0], cent[:, 1], cent[:, 2]
    rho = 1.0 + 0.3 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + amp * rng.uniform(-1, 1, n)
    v = np.stack([0.4 * np.sin(2 * np.pi * y), -0.3 * np.cos(2 * np.pi * z), 0.2 * np.sin(2 * np.pi * (x + z))])
    v = v + amp * rng.uniform(-1, 1, (3, n))
    p = 2.5 * (1 + 0.2 * np.cos(2 * np.pi * z) + amp * rng.uniform(-1, 1, n))
    u = np.empty((5, n))
    u[0], u[1:4], u[4] = rho, rho * v, p / 0.4 + 0.5 * rho * (v * v).sum(0)
    return np.ascontiguousarray(u.astype(dtype))
