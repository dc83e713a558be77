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

def hybrid_mesh(n=6, periodic=True, dtype=np.float64, seed=0, shuffle=False):
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

    seen = {}
    nbr, nrm, area, bnbr, bnrm, barea = [], [], [], [], [], []
    closed = np.zeros((len(elems), 3))
    for e, (o, faces, vf, cen) in enumerate(elems):
        for f in faces:
            a = geom(o, f, cen)
            closed[e] += a
            kf = key(o, f)
            if kf in seen:
                l, al = seen.pop(kf)
                assert np.allclose(al, -a, atol=1e-14), "non-matching interface"
                nbr += [l, e]
                A = np.linalg.norm(al)
                nrm += list(al / A)
                area.append(A)
            else:
                seen[kf] = (e, a)
    assert np.abs(closed).max() < 1e-14, "element surfaces are not closed"
    for kf, (e, a) in sorted(seen.items(), key=lambda t: t[1][0]):
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
    return conn, vol.astype(dtype), cent


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
