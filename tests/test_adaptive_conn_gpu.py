"""f-2, second half: connectivity of ADAPTIVE 2:1-balanced Cartesian forests built on the device from the leaf list
(t8b200_forest_connectivity) against the restatement of MeshManager::compute_connectivity_information over the
mini-forest (itself pinned bit for bit against the reference's own mesh manager, tests/test_reference_gpu.py): every
array equal, for every rank -- hanging faces, ghosts, x-faces, walls, 2-D and 3-D."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
KEYS = ("ranks", "indices", "face_neighbors", "face_normals", "face_areas", "x_face_neighbors", "x_face_normals",
        "x_face_areas")


def adapted(dim, level, periodic, rounds, P):
    f = oracle.Forest(dim, level, periodic)
    for rnd in range(rounds):
        lv, cent, vol, _ = f.elements()
        rng = np.random.default_rng(rnd)
        crit = np.where(np.abs(cent[:, dim - 1] - 0.45 + 0.1 * rnd) < 0.17, 20.0, 0.0) + rng.uniform(0, 1, len(lv))
        f = f.adapt(crit, 10.0, 1, level + 2, nranks=P)
    return f


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("dim,level,periodic,rounds,P", [(3, 3, True, 2, 1), (3, 3, True, 2, 3), (3, 2, False, 2, 2),
                                                          (2, 4, True, 3, 4), (2, 3, False, 2, 1), (3, 4, True, 0, 2)])
def test_device_connectivity_of_adaptive_forests_is_bit_exact(cuda, dim, level, periodic, rounds, P, dtype):
    import t8gpu_b200 as tb
    npdt = np.float64 if dtype == torch.float64 else np.float32
    f = adapted(dim, level, periodic, rounds, P)
    lv, cent, vol, _ = f.elements()
    if rounds:
        assert lv.min() < lv.max()                       # hanging faces are present
    keys = tb.morton_keys(dim, lv, cent)
    assert (np.diff(keys.astype(np.int64)) > 0).all()    # SFC order
    off = f.partition_offsets(P)
    for r in range(P):
        ref = f.connectivity(P, r, dtype=npdt)
        got = tb.forest_connectivity(dim, periodic, keys, lv, dtype, P, r, device=cuda)
        for k in ("n_local", "n_ghost", "n_faces", "n_bfaces", "n_xfaces"):
            assert int(got[k]) == int(ref[k]), (k, r)
        for k in KEYS:
            a, b = got[k].cpu().numpy(), ref[k]
            assert a.dtype == b.dtype and np.array_equal(a, b), (k, r)
        assert np.array_equal(got["volumes"].cpu().numpy(), vol[off[r]:off[r + 1]].astype(npdt))
        assert np.array_equal(got["centroids"].cpu().numpy().reshape(-1, 3), cent[off[r]:off[r + 1]].astype(npdt))


def test_device_connectivity_feeds_the_fused_path(cuda):
    """End to end without the host face loop: leaves -> device connectivity -> plan -> 3 steps == the oracle."""
    import t8gpu_b200 as tb
    from util import TOL, perturbed_kh, rel_linf
    f = adapted(3, 3, True, 2, 1)
    lv, cent, vol, _ = f.elements()
    conn = tb.forest_connectivity(3, True, tb.morton_keys(3, lv, cent), lv, torch.float64, device=cuda)
    sol = tb.EulerSolver(tb.conn_to_host(conn), conn["volumes"], torch.float64, device=cuda)
    u0, volh = perturbed_kh(f, np.float64, seed=4)
    sol.set_state(u0)
    ref_conn, u = f.connectivity(dtype=np.float64), u0
    dt = 0.05 * 2.0 ** -5
    for it in range(3):
        u, _, _ = oracle.iterate(ref_conn, volh, u, dt)
        sol.iterate(dt)
        assert rel_linf(sol.state().cpu().numpy(), u) <= (it + 1) * TOL[np.dtype(np.float64)]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("dim,level,periodic,rounds,P", [(3, 3, True, 2, 1), (3, 3, True, 2, 3), (3, 2, False, 2, 2),
                                                          (2, 4, True, 3, 4), (2, 3, False, 2, 1)])
def test_device_subgrid_connectivity_is_bit_exact(cuda, dim, level, periodic, rounds, P, dtype):
    """SubgridMeshManager layout (level differences, neighbour offsets, finer element second) from the same leaf list."""
    import t8gpu_b200 as tb
    npdt = np.float64 if dtype == torch.float64 else np.float32
    f = adapted(dim, level, periodic, rounds, P)
    lv, cent, vol, _ = f.elements()
    keys = tb.morton_keys(dim, lv, cent)
    seen_swapped = False
    for r in range(P):
        ref = f.connectivity(P, r, subgrid=True, dtype=npdt)
        got = tb.forest_connectivity(dim, periodic, keys, lv, dtype, P, r, device=cuda, subgrid=True)
        for k in ("n_local", "n_ghost", "n_faces", "n_bfaces", "n_xfaces"):
            assert int(got[k]) == int(ref[k]), (k, r)
        for k in KEYS + ("level_diff", "offsets", "x_level_diff", "x_offsets"):
            a, b = got[k].cpu().numpy(), ref[k]
            assert a.dtype == b.dtype and np.array_equal(a, b), (k, r)
        assert (ref["level_diff"] < 0).any()
        nb = ref["face_neighbors"][:2 * int(ref["n_faces"])].reshape(-1, 2)
        seen_swapped |= bool((nb[:, 0] >= int(ref["n_local"])).any())
    if P > 1 and periodic:
        assert seen_swapped          # a ghost in the first slot: the canonical swap of a finer ghost neighbour


def test_device_subgrid_connectivity_feeds_the_cell_plan(cuda):
    """leaves -> device subgrid connectivity -> cell-level plan -> 2 steps == the subgrid oracle."""
    import t8gpu_b200 as tb
    from util import TOL, rel_linf
    f = adapted(3, 2, True, 2, 1)
    lv, cent, vol, _ = f.elements()
    conn = tb.conn_to_host(tb.forest_connectivity(3, True, tb.morton_keys(3, lv, cent), lv, torch.float64, device=cuda,
                                                  subgrid=True))
    ref_conn = f.connectivity(subgrid=True, dtype=np.float64)
    sol = tb.SubgridEulerSolver(conn, vol, torch.float64, device=cuda, mode="fused")
    ref = tb.SubgridEulerSolver(ref_conn, vol, torch.float64, device=cuda, mode="fused")
    rng = np.random.default_rng(2)
    n = f.num_elements * 64
    u0 = np.stack([1 + 0.2 * rng.random(n), 0.1 * rng.standard_normal(n), 0.1 * rng.standard_normal(n),
                   0.1 * rng.standard_normal(n), 3 + 0.2 * rng.random(n)])
    sol.set_state(u0)
    ref.set_state(u0)
    for _ in range(2):
        sol.iterate(1e-3)
        ref.iterate(1e-3)
    assert torch.equal(sol.state(), ref.state())
