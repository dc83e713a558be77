"""GPU parity of the subgrid path: product (C ABI) vs the CPU oracle, the golden fixtures and the reference's own
kernels (oracle/_ref)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import ref_cuda
from util import TOL, rel_linf

pytestmark = pytest.mark.gpu
DT = {np.float64: torch.float64, np.float32: torch.float32}
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sg_forest(kind):
    if kind == "hex2":
        return oracle.Forest(3, 2), 2
    if kind == "quad3":
        return oracle.Forest(2, 3), 3
    if kind == "hex2_walls":
        return oracle.Forest(3, 2, periodic=False), 2
    if kind == "quad3_walls":
        return oracle.Forest(2, 3, periodic=False), 3
    if kind == "hex_amr":
        f = oracle.Forest(3, 2)
        lv, cent, vol, _ = f.elements()
        return f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < 0.2, 1.0, 0.0), 0.02, 1, 6), 3
    if kind == "quad_amr_walls":
        f = oracle.Forest(2, 3, periodic=False)
        lv, cent, vol, _ = f.elements()
        return f.adapt(np.where(np.abs(cent[:, 1] - 0.5) < 0.2, 1.0, 0.0), 0.02, 1, 6), 4
    raise KeyError(kind)


def sg_state(forest, dtype, seed):
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


def modes_for(dim):
    return ["unfused", "fused"]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["hex2", "quad3", "hex2_walls", "quad3_walls", "hex_amr", "quad_amr_walls"])
def test_subgrid_iterate_matches_oracle(cuda, kind, dtype):
    import t8gpu_b200
    forest, lvl = sg_forest(kind)
    conn = forest.connectivity(subgrid=True, dtype=dtype)
    u0, vol = sg_state(forest, dtype, seed=41)
    dt = 0.1 * 2.0 ** -(lvl + 2)
    for mode in modes_for(forest.dim):
        sol = t8gpu_b200.SubgridEulerSolver(conn, vol, DT[dtype], device=cuda, mode=mode)
        sol.set_state(u0)
        u = u0
        for it in range(6):
            u, _, _ = oracle.subgrid_iterate(conn, vol, u, dt)
            sol.iterate(dt)
            err = rel_linf(sol.state().cpu().numpy(), u)
            assert err <= (it + 1) * TOL[np.dtype(dtype)], (kind, mode, it, err)


@pytest.mark.parametrize("tag,dtype", [("f32", np.float32), ("f64", np.float64)])
@pytest.mark.parametrize("case", ["sg_hex2", "sg_quad3", "sg_hex2_amr", "sg_quad3_amr"])
def test_subgrid_product_vs_reference_golden(cuda, case, tag, dtype):
    """Product against the committed outputs of the reference's own subgrid kernels, on the reference's arrays."""
    import t8gpu_b200
    g = np.load(os.path.join(GOLD, case + "_" + tag + ".npz"))
    cnt = g["conn_counts"]
    conn = dict(dim=int(g["dim"]), n_local=int(cnt[0]), n_ghost=int(cnt[1]), n_faces=int(cnt[2]), n_bfaces=int(cnt[3]),
                face_neighbors=g["conn_face_neighbors"], face_normals=g["conn_face_normals"],
                face_areas=g["conn_face_areas"], level_diff=g["conn_level_diff"], offsets=g["conn_offsets"])
    for mode in modes_for(conn["dim"]):
        sol = t8gpu_b200.SubgridEulerSolver(conn, g["conn_volumes"], DT[dtype], device=cuda, mode=mode)
        sol.set_state(g["u0"])
        done = 0
        for k in g["snaps"]:
            for _ in range(int(k) - done):
                sol.iterate(float(g["dt"]))
            done = int(k)
            err = rel_linf(sol.state().cpu().numpy(), g["u_%d" % k])
            assert err <= done * TOL[np.dtype(dtype)], (case, mode, k, err)


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,level", [(3, 3), (2, 4)])
def test_subgrid_product_vs_reference_live(cuda, dim, level, dtype):
    """20 steps side by side with the reference's SubgridCompressibleEulerSolver, after the reference's own
    adapt + partition + connectivity rebuild."""
    import t8gpu_b200
    s = ref_cuda.RefSolver("sg", dtype, dim, level, True)
    f = oracle.Forest(dim, level, True)
    lv, cent, vol, _ = f.elements()
    crit = np.where(np.abs(cent[:, dim - 1] - 0.5) < 0.2, 1.0, 0.0).astype(dtype)
    s.mesh_adapt(crit)
    f = f.adapt(crit, 0.02, 1, 6)
    conn = s.connectivity()
    orc = f.connectivity(subgrid=True, dtype=dtype)
    for k in ("face_neighbors", "face_normals", "face_areas", "level_diff", "offsets"):
        assert np.array_equal(conn[k], orc[k]), k
    u0 = s.get_state()
    dt = 0.1 * 2.0 ** -(level + 3)
    for mode in modes_for(dim):
        s.set_state(u0)
        sol = t8gpu_b200.SubgridEulerSolver(conn, conn["volumes"], DT[dtype], device=cuda, mode=mode)
        sol.set_state(u0)
        for it in range(20):
            s.iterate(dt)
            sol.iterate(dt)
            err = rel_linf(sol.state().cpu().numpy(), s.get_state())
            assert err <= (it + 1) * TOL[np.dtype(dtype)], (mode, it, err)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind,P", [("hex2", 2), ("hex_amr", 3), ("quad_amr_walls", 2)])
def test_subgrid_multi_rank_on_one_gpu(cuda, kind, P, dtype):
    """P ranks of the fused subgrid stage emulated on one device: ghost cells through the (rank, index) tables,
    partition-boundary faces evaluated by both owners (x-faces), 2:1 hanging faces across ranks."""
    from multirank import MultiRankSubgrid
    forest, lvl = sg_forest(kind)
    conn1 = forest.connectivity(subgrid=True, dtype=dtype)
    u0, vol = sg_state(forest, dtype, seed=17)
    dt = 0.1 * 2.0 ** -(lvl + 2)
    mr = MultiRankSubgrid(forest, P, DT[dtype], cuda)
    mr.set_global_state(u0)
    u = u0
    for it in range(4):
        u, _, _ = oracle.subgrid_iterate(conn1, vol, u, dt)
        mr.iterate(dt)
        err = rel_linf(mr.global_state(), u)
        assert err <= (it + 1) * TOL[np.dtype(dtype)], (it, err)
