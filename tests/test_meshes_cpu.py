"""Host logic of the mixed-element workload (BASELINE config 5): the periodic tiling of the hex / prism / tet pattern and
its partition into per-rank arrays in the reference's layout, checked by brute force and through the CPU oracle (the
partitions of a mesh, evaluated with the owner-computes protocol, reproduce the one-rank step)."""
import numpy as np

import oracle
from t8gpu_b200 import meshes


def test_tiling_is_a_closed_conforming_mesh():
    c0, v0, x0, shift = meshes.hybrid_mesh(6, True, np.float64, with_shift=True)
    assert shift.shape == (c0["n_faces"], 3) and set(np.unique(shift)) <= {-1, 0, 1}
    big, vol, cent = meshes.tile_periodic_mesh(c0, v0, x0, shift, 2)
    assert big["n_local"] == 8 * c0["n_local"] and big["n_faces"] == 8 * c0["n_faces"] and abs(vol.sum() - 1) < 1e-12
    nb = big["face_neighbors"].reshape(-1, 2)
    a = big["face_normals"].reshape(-1, 3) * big["face_areas"][:, None]
    acc = np.zeros((big["n_local"], 3))
    np.add.at(acc, nb[:, 0], a)
    np.add.at(acc, nb[:, 1], -a)
    assert np.abs(acc).max() < 1e-15                       # every element's surface is closed
    d = cent[nb[:, 1]] - cent[nb[:, 0]]
    d -= np.rint(d)
    assert np.abs(d).max() <= 1.0 / 12 + 1e-12             # face neighbours are spatial neighbours (periodic distance)
    assert (np.einsum("ij,ij->i", d, big["face_normals"].reshape(-1, 3)) > 0).all()   # normals point left -> right
    # the tiling of the pattern IS the pattern generated at twice the size, as a set of elements
    c1, v1, x1 = meshes.hybrid_mesh(12, True, np.float64)
    assert c1["n_local"] == big["n_local"] and c1["n_faces"] == big["n_faces"]
    assert np.allclose(np.sort(v1), np.sort(vol)) and np.allclose(np.sort(x1.sum(1)), np.sort(cent.sum(1)))


def test_partitions_reproduce_the_one_rank_step():
    c0, v0, x0, shift = meshes.hybrid_mesh(6, True, np.float64, with_shift=True)
    big, vol, cent = meshes.tile_periodic_mesh(c0, v0, x0, shift, 2)
    u0 = meshes.smooth_state(cent, np.float64, seed=3)
    dt = 0.02 / 12
    ref, _, _ = oracle.iterate(big, vol, u0, dt)
    P = 3
    parts = [meshes.partition_flat_mesh(big, vol, P, r) for r in range(P)]
    off = parts[0][0]["offsets_global"]
    seen = set()
    for r, (conn, lvol) in enumerate(parts):
        nl, ng = conn["n_local"], conn["n_ghost"]
        assert nl == off[r + 1] - off[r] and np.array_equal(lvol, vol[off[r]:off[r + 1]])
        assert (conn["ranks"][:nl] == r).all() and np.array_equal(conn["indices"][:nl], np.arange(nl))
        assert (conn["ranks"][nl:] != r).all()
        glob = np.concatenate([np.arange(off[r], off[r + 1]), off[conn["ranks"][nl:]] + conn["indices"][nl:]])
        assert len(set(glob.tolist())) == nl + ng           # every ghost once
        for tag, nbr in (("m", conn["face_neighbors"][:2 * conn["n_faces"]]), ("x", conn["x_face_neighbors"])):
            for a, b in nbr.reshape(-1, 2):
                assert a < nl                                # seen from a local element
                if b >= nl:                                  # lower rank owns the face, the higher rank mirrors it
                    assert (conn["ranks"][b] > r) == (tag == "m")
                seen.add((r, min(int(glob[a]), int(glob[b])), max(int(glob[a]), int(glob[b]))))
    allf = set((min(a, b), max(a, b)) for a, b in big["face_neighbors"].reshape(-1, 2).tolist())
    assert set((a, b) for _, a, b in seen) == allf           # every face of the mesh is evaluated somewhere

    # owner computes with the oracle arithmetic: regular + x faces of every rank, local accumulators only
    def stage(k, cur, prev):
        out = []
        for r, (conn, lvol) in enumerate(parts):
            nl, ng = conn["n_local"], conn["n_ghost"]
            ext = np.zeros((5, nl + ng))
            ext[:, :nl] = cur[r]
            for g in range(ng):
                ext[:, nl + g] = cur[int(conn["ranks"][nl + g])][:, int(conn["indices"][nl + g])]
            flux = np.zeros_like(ext)
            oracle.flux_faces(conn, np.ascontiguousarray(ext), flux, np.zeros(max(1, conn["n_faces"])))
            if conn["n_xfaces"]:
                xc = dict(n_faces=conn["n_xfaces"], n_bfaces=0, face_neighbors=conn["x_face_neighbors"],
                          face_normals=conn["x_face_normals"], face_areas=conn["x_face_areas"])
                oracle.flux_faces(xc, np.ascontiguousarray(ext), flux, np.zeros(conn["n_xfaces"]))
            o = np.zeros_like(cur[r])
            oracle.rk_stage(k, prev[r], cur[r], o, np.ascontiguousarray(flux[:, :nl]), np.ascontiguousarray(lvol), dt)
            out.append(o)
        return out

    prev = [np.ascontiguousarray(u0[:, off[r]:off[r + 1]]) for r in range(P)]
    s1 = stage(1, prev, prev)
    s2 = stage(2, s1, prev)
    nxt = stage(3, s2, prev)
    got = np.concatenate(nxt, axis=1)
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()


def test_send_csr_groups_the_send_list_by_source_element():
    """multi.send_csr: (src, dst_rank, dst_idx) sorted by destination -> CSR by source element, stable."""
    import torch
    from t8gpu_b200 import multi
    src = torch.tensor([5, 2, 5, 0, 2, 7], dtype=torch.int32)
    drk = torch.tensor([1, 1, 2, 2, 3, 3], dtype=torch.int32)
    dix = torch.tensor([10, 11, 20, 21, 30, 31], dtype=torch.int32)
    off, rk, ix = multi.send_csr((src, drk, dix), 8, "cpu")
    assert off.tolist() == [0, 1, 1, 3, 3, 3, 5, 5, 6]
    assert rk.tolist() == [2, 1, 3, 1, 2, 3] and ix.tolist() == [21, 11, 30, 10, 20, 31]
    off, rk, ix = multi.send_csr((src[:0], drk[:0], dix[:0]), 3, "cpu")
    assert off.tolist() == [0, 0, 0, 0]
