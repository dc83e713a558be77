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


def test_bench_subgrid_connectivity_and_cells(cuda):
    """bench.py --workload subgrid builds its mesh without the oracle: the arrays must be the ones the reference's
    SubgridMeshManager produces for the same forest (bit-exact), and the cell centres those of the reference's IC."""
    import oracle
    import t8gpu_b200 as tb
    from bench_subgrid import cell_centers, subgrid_connectivity
    L = 3
    f = oracle.Forest(3, L)
    ref = f.connectivity(subgrid=True, dtype=np.float64)
    got = tb.conn_to_host(subgrid_connectivity(L, torch.float64, 0, 1, cuda, (1, 1, 1)))
    for key in ("face_neighbors", "face_normals", "face_areas", "level_diff", "offsets"):
        assert np.array_equal(got[key], ref[key]), key
    lv, cent, vol, _ = f.elements()
    assert np.array_equal(got["volumes"], vol)
    u = [torch.zeros(f.num_elements * 64, dtype=torch.float64, device=cuda) for _ in range(5)]
    tb.init_kelvin_helmholtz(3, cell_centers(torch.as_tensor(cent).to(cuda), L, torch.float64), u)
    want = oracle.subgrid_init_kh(3, cent, lv, np.float64)
    assert rel_linf(torch.stack(u).cpu().numpy(), want) <= 1e-14


def test_bench_subgrid_brick_ranks_agree_with_one_rank(cuda):
    """The weak-scaling subgrid mesh: 2 ranks (one tree each, emulated on one device) vs the same brick on one rank."""
    import t8gpu_b200 as tb
    from bench_subgrid import cell_centers, subgrid_connectivity
    L, P, brick, dt = 2, 2, (2, 1, 1), 0.1 * 2.0 ** -4
    c1 = subgrid_connectivity(L, torch.float64, 0, 1, cuda, brick)
    one = tb.SubgridEulerSolver(dict(tb.conn_to_host(c1), dim=3), c1["volumes"].cpu().numpy(), torch.float64,
                                device=cuda, mode="fused")
    tb.init_kelvin_helmholtz(3, cell_centers(c1["centroids"], L, torch.float64), one.variables(one.next))
    rng = np.random.default_rng(4)
    u0 = one.state().cpu().numpy() * (1 + 0.02 * rng.uniform(-1, 1, (5, one.nc)))
    one.set_state(u0)
    conns = [subgrid_connectivity(L, torch.float64, r, P, cuda, brick) for r in range(P)]
    nc = [int(c["n_local"]) * 64 for c in conns]
    off = np.concatenate([[0], np.cumsum(nc)])
    assert off[-1] == one.nc and all(int(c["n_xfaces"]) + int(c["n_ghost"]) > 0 for c in conns)
    bufs = [torch.zeros((25, nc[r]), dtype=torch.float64, device=cuda) for r in range(P)]
    for r in range(P):
        bufs[r][0:5] = torch.as_tensor(u0[:, off[r]:off[r + 1]]).to(cuda)
    tabs = {s: tb.RankTables([[bufs[r][s * 5 + k] for k in range(5)] for r in range(P)], cuda) for s in range(5)}
    plans = [tb.SubgridPlan(tb.conn_to_host(c), c["volumes"].cpu().numpy(), torch.float64) for c in conns]
    nxt, prv = 0, 3
    for it in range(3):
        one.iterate(dt)
        nxt, prv = prv, nxt
        for stage, sin, sout in ((1, prv, 1), (2, 1, 2), (3, 2, nxt)):
            for r in range(P):
                v = lambda s: [bufs[r][s * 5 + k] for k in range(5)]  # noqa: E731
                plans[r].stage(stage, v(sin), v(prv), v(sout), conns[r]["volumes"], dt, in_all=tabs[sin])
        got = np.concatenate([bufs[r][nxt * 5:(nxt + 1) * 5].cpu().numpy() for r in range(P)], axis=1)
        assert rel_linf(got, one.state().cpu().numpy()) <= 1e-12


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("dim", [2, 3])
def test_z_order_output_kernel(cuda, dim, dtype):
    """t8b200_subgrid_z_order: out[e*S + morton(i,j,k)] = double(in[e*S + i + 4j + 16k])
    (column_major_to_z_order, t8gpu/mesh/subgrid_mesh_manager.inl:1007-1049, plus the host widening loop :1106-1109)."""
    import t8gpu_b200
    S = 64 if dim == 3 else 16
    n = 1000
    x = torch.rand(n * S, dtype=dtype, device=cuda)
    got = t8gpu_b200.subgrid_z_order(dim, x).cpu().numpy()
    perm = np.zeros(S, np.int64)
    for flat in range(S):
        i, j, k = flat & 3, (flat >> 2) & 3, flat >> 4
        mo = 0
        for l in range(2):
            mo |= ((i >> l) & 1) << (dim * l) | ((j >> l) & 1) << (dim * l + 1) | (((k >> l) & 1) << (dim * l + 2) if dim == 3 else 0)
        perm[mo] = flat
    want = x.cpu().numpy().astype(np.float64).reshape(n, S)[:, perm].reshape(-1)
    assert got.dtype == np.float64 and np.array_equal(got, want)
