"""GPU parity: the CUDA product (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

import oracle
from util import TOL, perturbed_kh, rel_linf

pytestmark = pytest.mark.gpu

DT = {np.float64: torch.float64, np.float32: torch.float32}


def make_forest(kind):
    if kind == "quad6":      # BASELINE configs[0]: periodic quad level 6, MeshManager<...,3> normals
        return oracle.Forest(2, 6), 6
    if kind == "hex4":
        return oracle.Forest(3, 4), 4
    if kind == "hex3_walls":  # non-periodic: reflective boundary faces
        return oracle.Forest(3, 3, periodic=False), 3
    if kind == "hex_amr":    # hanging faces (2:1)
        f = oracle.Forest(3, 3)
        lv, cent, vol, _ = f.elements()
        crit = np.where(np.abs(cent[:, 2] - 0.5) < 0.2, 20.0, 0.0)
        return f.adapt(crit, 10.0, 1, 4), 4
    if kind == "quad_amr_walls":
        f = oracle.Forest(2, 4, periodic=False)
        lv, cent, vol, _ = f.elements()
        crit = np.where(np.abs(cent[:, 1] - 0.5) < 0.2, 20.0, 0.0)
        return f.adapt(crit, 10.0, 1, 6), 5
    raise KeyError(kind)


def rotated(conn, seed=7):
    """Same topology, but the mesh is rotated in space (general unit normals) and the face areas are perturbed:
    exercises the uncompressed-geometry path of the tile plan."""
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    c = dict(conn)
    dt = conn["face_normals"].dtype
    n = conn["face_normals"].reshape(-1, 3).astype(np.float64) @ q.T
    c["face_normals"] = np.ascontiguousarray(n.reshape(-1).astype(dt))
    c["face_areas"] = (conn["face_areas"].astype(np.float64) * rng.uniform(0.8, 1.2, conn["face_areas"].shape)).astype(dt)
    return c


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("mode", ["fused", "unfused"])
@pytest.mark.parametrize("kind", ["hex3_walls", "hex_amr"])
def test_general_normals(cuda, kind, mode, dtype):
    import t8gpu_b200
    forest, lvl = make_forest(kind)
    conn = rotated(forest.connectivity(dtype=dtype))
    u0, vol = perturbed_kh(forest, dtype, seed=11)
    dt = 0.05 * 2.0 ** -lvl
    sol = t8gpu_b200.EulerSolver(conn, vol, DT[dtype], device=cuda, mode=mode)
    sol.set_state(u0)
    u = u0
    for it in range(5):
        u, _, _ = oracle.iterate(conn, vol, u, dt)
        sol.iterate(dt)
        err = rel_linf(sol.state().cpu().numpy(), u)
        assert err <= (it + 1) * TOL[np.dtype(dtype)], (kind, mode, it, err)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("mode", ["fused", "unfused"])
@pytest.mark.parametrize("kind", ["quad6", "hex4", "hex3_walls", "hex_amr", "quad_amr_walls"])
def test_iterate_matches_oracle(cuda, kind, mode, dtype):
    import t8gpu_b200
    forest, lvl = make_forest(kind)
    conn = forest.connectivity(dtype=dtype)
    u0, vol = perturbed_kh(forest, dtype, seed=1)
    dt = 0.1 * 2.0 ** -lvl
    sol = t8gpu_b200.EulerSolver(conn, vol, DT[dtype], device=cuda, mode=mode)
    sol.set_state(u0)
    u = u0
    nsteps = 10
    for it in range(nsteps):
        sp = np.zeros(conn["n_faces"] + conn["n_bfaces"], dtype)
        u, s1, s2 = oracle.iterate(conn, vol, u, dt, speed=sp)
        sol.iterate(dt)
        got = sol.state().cpu().numpy()
        err = rel_linf(got, u)
        # per-step tolerance of the north star; both sides advance from their own previous state, so allow the
        # accumulated bound it * tol
        assert err <= (it + 1) * TOL[np.dtype(dtype)], (kind, mode, it, err)
        vmax = float(sol.max_wave_speed().item())
        assert abs(vmax - oracle.max_speed(sp)) <= 4 * np.finfo(dtype).eps * vmax * (10 if dtype == np.float32 else 1e3)
    dt_ref = oracle.compute_timestep(sp, dtype(0.7), 4)
    assert abs(sol.compute_timestep() - dt_ref) <= 1e-5 * dt_ref


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_single_step_from_identical_state(cuda, dtype):
    """One step from bit-identical input: the pure per-step error (no accumulation)."""
    import t8gpu_b200
    forest, lvl = make_forest("hex_amr")
    conn = forest.connectivity(dtype=dtype)
    u0, vol = perturbed_kh(forest, dtype, seed=2)
    dt = 0.1 * 2.0 ** -lvl
    ref, _, _ = oracle.iterate(conn, vol, u0, dt)
    for mode in ("fused", "unfused"):
        sol = t8gpu_b200.EulerSolver(conn, vol, DT[dtype], device=cuda, mode=mode)
        sol.set_state(u0)
        sol.iterate(dt)
        assert rel_linf(sol.state().cpu().numpy(), ref) <= TOL[np.dtype(dtype)]


def test_hundred_steps_fp64(cuda):
    """north_star: within 1e-12 relative L-infinity per step over 100 steps (fp64), here on the level-6 quad mesh of
    configs[0] with the unperturbed Kelvin-Helmholtz data and the reference's dt = 0.1 * 2^-6."""
    import t8gpu_b200
    forest = oracle.Forest(2, 6)
    conn = forest.connectivity(dtype=np.float64)
    lv, cent, vol, _ = forest.elements()
    u = oracle.init_kh_points(2, cent, np.float64)
    sol = t8gpu_b200.EulerSolver(conn, vol, torch.float64, device=cuda, mode="fused")
    sol.set_state(u)
    dt = 0.1 * 2.0 ** -6
    worst = 0.0
    for it in range(100):
        # restart the oracle from the product's state each step: measures the per-step difference
        cur = sol.state().cpu().numpy().copy()
        ref, _, _ = oracle.iterate(conn, vol, cur, dt)
        sol.iterate(dt)
        worst = max(worst, rel_linf(sol.state().cpu().numpy(), ref))
    assert worst <= 1e-12, worst


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,level,P", [(2, 6, 1), (3, 4, 1), (3, 4, 2), (3, 3, 3), (2, 5, 4), (3, 1, 1), (3, 0, 1)])
def test_device_cartesian_connectivity_bit_exact(cuda, dim, level, P, dtype):
    """The device-built connectivity equals, bit for bit, the mini-forest restatement of
    MeshManager::compute_connectivity_information for every rank."""
    import t8gpu_b200
    forest = oracle.Forest(dim, level)
    lv, cent, vol, _ = forest.elements()
    off = forest.partition_offsets(P)
    for r in range(P):
        ref = forest.connectivity(P, r, dtype=dtype)
        got = t8gpu_b200.conn_to_host(t8gpu_b200.cartesian_uniform_connectivity(dim, level, DT[dtype], P, r))
        for key in ("n_local", "n_ghost", "n_faces", "n_bfaces", "n_xfaces"):
            assert got[key] == ref[key], (key, r)
        for key in ("ranks", "indices", "face_neighbors", "face_normals", "face_areas", "x_face_neighbors",
                    "x_face_normals", "x_face_areas"):
            assert got[key].dtype == ref[key].dtype, key
            assert np.array_equal(got[key], ref[key]), (key, r)
        assert np.array_equal(got["volumes"], vol[off[r]:off[r + 1]].astype(dtype))
        assert np.array_equal(got["centroids"].reshape(-1, 3), cent[off[r]:off[r + 1]].astype(dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_kelvin_helmholtz_initial_state(cuda, dtype):
    import t8gpu_b200
    for dim, level in ((2, 5), (3, 3)):
        forest = oracle.Forest(dim, level)
        lv, cent, vol, _ = forest.elements()
        c = torch.as_tensor(cent.astype(dtype)).to(cuda).contiguous()
        u = [torch.empty(len(cent), dtype=DT[dtype], device=cuda) for _ in range(5)]
        t8gpu_b200.init_kelvin_helmholtz(dim, c, u)
        got = torch.stack(u).cpu().numpy()
        ref = oracle.init_kh_points(dim, cent.astype(dtype), dtype)
        assert rel_linf(got, ref) <= 4 * np.finfo(dtype).eps


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind,P", [("hex4", 2), ("hex_amr", 3), ("quad6", 4)])
@pytest.mark.parametrize("mode", ["fused", "unfused"])
def test_multi_rank_on_one_gpu(cuda, kind, P, mode, dtype):
    """P ranks emulated on one device (no kernel waits on another: stage s of every rank is launched before stage
    s+1 of any).  Ghost reads go through the (rank, index) tables exactly as across GPUs."""
    from multirank import MultiRankEuler
    forest, lvl = make_forest(kind)
    conn1 = forest.connectivity(dtype=dtype)
    u0, vol = perturbed_kh(forest, dtype, seed=3)
    dt = 0.1 * 2.0 ** -lvl
    mr = MultiRankEuler(forest, P, DT[dtype], cuda, mode=mode)
    mr.set_global_state(u0)
    u = u0
    for it in range(5):
        u, _, _ = oracle.iterate(conn1, vol, u, dt)
        mr.iterate(dt)
        err = rel_linf(mr.global_state(), u)
        assert err <= (it + 1) * TOL[np.dtype(dtype)], (it, err)


def test_ragged_and_empty_inputs(cuda):
    import t8gpu_b200
    # element count not a multiple of the chunk size, and a mesh smaller than one chunk
    for kind in ("hex3_walls", "quad_amr_walls"):
        forest, lvl = make_forest(kind)
        assert forest.num_elements % 256 != 0 or kind == "hex3_walls"
    forest = oracle.Forest(3, 1)   # 8 elements, every face pair appears twice (periodic, 2 cells per axis)
    conn = forest.connectivity(dtype=np.float64)
    u0, vol = perturbed_kh(forest, np.float64, seed=4)
    ref, _, _ = oracle.iterate(conn, vol, u0, 0.01)
    sol = t8gpu_b200.EulerSolver(conn, vol, torch.float64, device=cuda)
    sol.set_state(u0)
    sol.iterate(0.01)
    assert rel_linf(sol.state().cpu().numpy(), ref) <= 1e-12
    # level 0: one element, no faces at all -> state only rescaled by the RK coefficients
    forest = oracle.Forest(3, 0)
    conn = forest.connectivity(dtype=np.float64)
    assert conn["n_faces"] == 0
    u0, vol = perturbed_kh(forest, np.float64, seed=5)
    ref, _, _ = oracle.iterate(conn, vol, u0, 0.01)
    for mode in ("fused", "unfused"):
        sol = t8gpu_b200.EulerSolver(conn, vol, torch.float64, device=cuda, mode=mode)
        sol.set_state(u0)
        sol.iterate(0.01)
        assert rel_linf(sol.state().cpu().numpy(), ref) <= 1e-12


def test_brick_2x2x2_equals_cube_one_level_up(cuda):
    """A (2,2,2) brick of level-L trees is the level-(L+1) cube in the same element order: checks the multi-tree
    neighbour arithmetic of the device builder against the mini-forest, rank by rank."""
    import t8gpu_b200
    L = 3
    forest = oracle.Forest(3, L + 1)
    for P in (1, 8):
        for r in range(P):
            ref = forest.connectivity(P, r, dtype=np.float64)
            got = t8gpu_b200.conn_to_host(t8gpu_b200.cartesian_uniform_connectivity(3, L, torch.float64, P, r,
                                                                                      brick=(2, 2, 2)))
            for key in ("ranks", "indices", "face_neighbors", "face_normals", "x_face_neighbors", "x_face_normals"):
                assert np.array_equal(got[key], ref[key]), (key, P, r)
            # geometry differs by the unit: brick trees are unit cubes, the cube's octants have edge 1/2
            assert np.array_equal(got["face_areas"], ref["face_areas"] * 4)


@pytest.mark.parametrize("brick,P", [((2, 1, 1), 2), ((2, 2, 1), 4)])
def test_brick_partitions_agree_with_single_rank(cuda, brick, P):
    """Weak-scaling meshes: P ranks (one tree each, emulated on one device) vs the same brick on one rank."""
    import t8gpu_b200
    from t8gpu_b200.solver import NB_STEPS, NVAR
    L, dt = 3, 0.1 * 2.0 ** -3
    c1 = t8gpu_b200.cartesian_uniform_connectivity(3, L, torch.float64, 1, 0, brick=brick)
    n1 = int(c1["n_local"])
    one = t8gpu_b200.EulerSolver(t8gpu_b200.conn_to_host(c1), c1["volumes"].cpu().numpy(), torch.float64, device=cuda)
    t8gpu_b200.init_kelvin_helmholtz(3, c1["centroids"], one.variables(one.next))
    rng = np.random.default_rng(9)
    u0 = one.state().cpu().numpy() * (1 + 0.02 * rng.uniform(-1, 1, (5, n1)))
    one.set_state(u0)
    conns = [t8gpu_b200.cartesian_uniform_connectivity(3, L, torch.float64, P, r, brick=brick) for r in range(P)]
    ns = [int(c["n_local"]) for c in conns]
    off = np.concatenate([[0], np.cumsum(ns)])
    bufs = []
    for r in range(P):
        b = torch.zeros((NVAR * NB_STEPS + 1, ns[r]), dtype=torch.float64, device=cuda)
        b[NVAR * NB_STEPS] = conns[r]["volumes"]
        b[0:5] = torch.as_tensor(u0[:, off[r]:off[r + 1]]).to(cuda)
        bufs.append(b)
    tabs = {s: t8gpu_b200.RankTables([[bufs[r][s * NVAR + k] for k in range(NVAR)] for r in range(P)], cuda)
            for s in range(NB_STEPS)}
    plans = [t8gpu_b200.Plan(t8gpu_b200.conn_to_host(c), torch.float64) for c in conns]
    nxt, prv = 0, 3
    for it in range(4):
        one.iterate(dt)
        nxt, prv = prv, nxt
        for stage, sin, sout in ((1, prv, 1), (2, 1, 2), (3, 2, nxt)):
            for r in range(P):
                v = lambda s: [bufs[r][s * NVAR + k] for k in range(NVAR)]  # noqa: E731
                plans[r].stage(stage, v(sin), v(prv), v(sout), bufs[r][NVAR * NB_STEPS], dt, in_all=tabs[sin])
        got = np.concatenate([bufs[r][nxt * NVAR:(nxt + 1) * NVAR].cpu().numpy() for r in range(P)], axis=1)
        assert rel_linf(got, one.state().cpu().numpy()) <= 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind,general", [("hex4", False), ("hex_amr", False), ("hex_amr", True), ("quad_amr_walls", False)])
def test_split_chunks(cuda, kind, general, dtype, monkeypatch):
    """Blocks of 256 elements whose halo does not fit the kernel's shared memory are split into smaller chunks
    (forced here by lowering the limit): same results, more chunks."""
    import t8gpu_b200
    forest, lvl = make_forest(kind)
    conn = forest.connectivity(dtype=dtype)
    if general:
        conn = rotated(conn)
    u0, vol = perturbed_kh(forest, dtype, seed=5)
    dt = 0.05 * 2.0 ** -lvl
    whole = t8gpu_b200.EulerSolver(conn, vol, DT[dtype], device=cuda, mode="fused")
    limit = 40 if forest.dim == 3 else 12
    monkeypatch.setenv("T8B200_TEST_MAX_HALO", str(limit))
    sol = t8gpu_b200.EulerSolver(conn, vol, DT[dtype], device=cuda, mode="fused")
    monkeypatch.delenv("T8B200_TEST_MAX_HALO")
    assert sol.plan.info["n_chunks"] > whole.plan.info["n_chunks"]
    assert sol.plan.info["max_halo"] <= limit
    sol.set_state(u0)
    u = u0
    for it in range(3):
        u, _, _ = oracle.iterate(conn, vol, u, dt)
        sol.iterate(dt)
        err = rel_linf(sol.state().cpu().numpy(), u)
        assert err <= (it + 1) * TOL[np.dtype(dtype)], (kind, it, err)
    assert float(sol.max_wave_speed().item()) > 0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("mode", ["fused", "unfused"])
@pytest.mark.parametrize("periodic,shuffle", [(True, False), (False, False), (True, True)])
def test_hybrid_tet_prism_hex_mesh(cuda, periodic, shuffle, mode, dtype):
    """Mixed element types (BASELINE config 5 at kernel level): hexahedra (8 faces after splitting their horizontal
    quads), prisms (5) and tetrahedra (4) with general normals, optionally in a scrambled element order (no spatial
    locality at all: large halos, blocks get split)."""
    import t8gpu_b200
    from util import hybrid_mesh, smooth_state
    conn, vol, cent = hybrid_mesh(6, periodic, dtype, shuffle=shuffle)
    u0 = smooth_state(cent, dtype, seed=31)
    dt = 0.02 / 6
    sol = t8gpu_b200.EulerSolver(conn, vol, DT[dtype], device=cuda, mode=mode)
    if mode == "fused" and shuffle:
        assert sol.plan.info["n_chunks"] > (conn["n_local"] + 255) // 256
    sol.set_state(u0)
    u = u0
    for it in range(4):
        u, _, _ = oracle.iterate(conn, vol, u, dt)
        sol.iterate(dt)
        err = rel_linf(sol.state().cpu().numpy(), u)
        assert err <= (it + 1) * TOL[np.dtype(dtype)], (it, err)
    # conservation: sum(vol * u) is preserved on the periodic mesh (fluxes cancel pairwise)
    if periodic and dtype == np.float64:
        tot0 = (u0.astype(np.float64) * vol).sum(1)
        tot1 = (sol.state().cpu().numpy() * vol).sum(1)
        assert np.abs(tot1 - tot0).max() <= 1e-12 * np.abs(tot0).max()


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("brick,P,L", [((2, 1, 1), 2, 3), ((2, 2, 1), 4, 3), ((2, 1, 1), 2, 4)])
def test_self_ordering_stage_kernels_bitwise(cuda, brick, P, L, dtype):
    """t8b200_fused_stage_sync_*: the stage kernels order themselves through the mailboxes (partition-boundary chunks
    first: wait for the peers' previous launch, the last one signals) and read dt from device memory.  P ranks emulated
    on one device, stage after stage (so no kernel ever waits for a later one).  Owner-computes is deterministic: the
    result must be BITWISE the one-rank run of the same brick, and bitwise the run ordered by plain launches."""
    import t8gpu_b200 as tb
    from t8gpu_b200.solver import NB_STEPS, NVAR
    # level 3: generic chunks (a tree 8 elements across is its own neighbour); level 4: structured chunks
    dt = 0.1 * 2.0 ** -L
    c1 = tb.cartesian_uniform_connectivity(3, L, dtype, 1, 0, brick=brick)
    n1 = int(c1["n_local"])
    one = tb.EulerSolver(tb.conn_to_host(c1), c1["volumes"].cpu().numpy(), dtype, device=cuda)
    tb.init_kelvin_helmholtz(3, c1["centroids"], one.variables(one.next))
    rng = np.random.default_rng(9)
    u0 = (one.state().cpu().numpy() * (1 + 0.02 * rng.uniform(-1, 1, (5, n1)))).astype(one.state().cpu().numpy().dtype)
    one.set_state(u0)
    conns = [tb.cartesian_uniform_connectivity(3, L, dtype, P, r, brick=brick) for r in range(P)]
    ns = [int(c["n_local"]) for c in conns]
    off = np.concatenate([[0], np.cumsum(ns)])
    plans = [tb.Plan(tb.conn_to_host(c), dtype) for c in conns]
    tail_plans = [tb.Plan(tb.conn_to_host(c), dtype, ghost_tail=True) for c in conns]
    mails = [tb.PeerMailboxes(r, P, cuda) for r in range(P)]
    for m in mails:
        m.set_table([x.buf.ptr for x in mails])     # same device: the mailboxes are plain device pointers
    dt_dev = torch.tensor([dt], dtype=dtype, device=cuda)

    def run(mode):
        """mode: "plain" (ordered by the launch sequence, direct ghost reads), "kernel" (self-ordering stage kernels, dt
        from device memory), "tail" (ghost tail: pull kernel, then stage kernels that read no other rank's rows)."""
        pl = tail_plans if mode == "tail" else plans
        cols = [ns[r] + pl[r].n_tail for r in range(P)]
        if mode == "tail":
            assert all(pl[r].n_tail == int(conns[r]["n_ghost"]) for r in range(P))
        bufs = []
        for r in range(P):
            b = torch.zeros((NVAR * NB_STEPS + 1, cols[r]), dtype=dtype, device=cuda)
            b[NVAR * NB_STEPS, :ns[r]] = conns[r]["volumes"]
            b[0:5, :ns[r]] = torch.as_tensor(u0[:, off[r]:off[r + 1]]).to(cuda)
            bufs.append(b)
        tabs = {s: tb.RankTables([[bufs[r][s * NVAR + k] for k in range(NVAR)] for r in range(P)], cuda)
                for s in range(NB_STEPS)}
        vmax = [torch.zeros(1, dtype=dtype, device=cuda) for _ in range(P)]
        if mode == "kernel":
            # "state written by other means": a t8b200_peer_barrier starts the epoch sequence on real GPUs; the ranks
            # emulated here share one stream, where a barrier kernel would wait for kernels behind it, so its effect
            # (every stage slot of parity e & 1 carries epoch e) is written directly
            for m in mails:
                e = m.stage_epoch + 1
                slots = m.buf.tensor((4 * P, 2), torch.int64)       # (value, epoch) per slot
                slots[(e & 1) * P:(e & 1) * P + P, 1] = e
                m.stage_epoch = e
        nxt, prv = 0, 3
        for it in range(4):
            nxt, prv = prv, nxt
            for stage, sin, sout in ((1, prv, 1), (2, 1, 2), (3, 2, nxt)):
                if mode == "tail":
                    for r in range(P):
                        pl[r].pull([bufs[r][sin * NVAR + k] for k in range(NVAR)], tabs[sin])
                for r in range(P):
                    v = lambda s: [bufs[r][s * NVAR + k] for k in range(NVAR)]  # noqa: E731
                    kw = dict(dt_dev=dt_dev, sync=mails[r]) if mode == "kernel" else {}
                    pl[r].stage(stage, v(sin), v(prv), v(sout), bufs[r][NVAR * NB_STEPS], 0.0 if kw else dt,
                                in_all=None if mode == "tail" else tabs[sin],
                                speed_max=vmax[r] if stage == 3 else None, **kw)
        torch.cuda.synchronize()
        return (np.concatenate([bufs[r][nxt * NVAR:(nxt + 1) * NVAR, :ns[r]].cpu().numpy() for r in range(P)], axis=1),
                max(float(x.item()) for x in vmax))

    plain, vm_plain = run("plain")
    own, vm_own = run("kernel")
    tail, vm_tail = run("tail")
    assert np.array_equal(tail, plain) and vm_tail == vm_plain
    for it in range(4):
        one.iterate(dt)
    ref = one.state().cpu().numpy()
    assert np.array_equal(own, plain) and vm_own == vm_plain
    assert np.array_equal(own, ref) and vm_own == float(one.max_wave_speed().item())
    for p in plans:
        assert p.info["n_chunks"] > 0
    for m in mails:
        assert m.stage_epoch == 1 + 12 and int(m.counter.item()) == 0   # one barrier + 4 steps x 3 stages
        m.close()


def test_timestep_kernel_and_alias_rejection(cuda):
    """t8b200_timestep_*: cfl * length / vmax capped by dt_cap, on the device (solver.cu:225-228); the fused stage
    rejects an output that aliases an input."""
    import ctypes as C
    import t8gpu_b200 as tb
    for dt_, npdt in ((torch.float64, np.float64), (torch.float32, np.float32)):
        vmax = torch.tensor([3.5], dtype=dt_, device=cuda)
        out = torch.zeros(1, dtype=dt_, device=cuda)
        tb.timestep(vmax, 0.7, 0.5 ** 4, 0.0, out)
        assert float(out.item()) == float(npdt(0.7) * npdt(0.5 ** 4) / npdt(3.5))
        tb.timestep(vmax, 0.7, 0.5 ** 4, 1e-3, out)
        assert float(out.item()) == float(npdt(1e-3))
        vmax.zero_()
        tb.timestep(vmax, 0.7, 0.5 ** 4, 1e-3, out)      # vmax == 0: the cap
        assert float(out.item()) == float(npdt(1e-3))
    f = oracle.Forest(3, 2)
    conn = f.connectivity(dtype=np.float64)
    sol = tb.EulerSolver(conn, f.elements()[2], torch.float64, device=cuda)
    v = sol.variables(0)
    with pytest.raises(tb.CudaError):
        sol.plan.stage(1, v, None, v, sol.volume(), 1e-3)


def test_peer_barrier_single_rank(cuda):
    """One rank: the mailbox barrier passes the value through and accepts increasing epochs (the multi-GPU semantics are
    exercised by bench.py under torchrun: kernels of different ranks must run on different GPUs)."""
    import ctypes as C
    import t8gpu_b200 as tb
    mb = tb.PeerMailboxes(0, 1, cuda)
    mb.exchange([mb.handle])
    for dt in (torch.float64, torch.float32):
        v = torch.tensor([3.25], dtype=dt, device=cuda)
        out = torch.zeros(1, dtype=dt, device=cuda)
        mb.barrier(v, out)
        mb.barrier()
        torch.cuda.synchronize()
        assert float(out[0]) == 3.25
    # push + barrier in one launch: one rank pushing into its own rows (many CTAs: the last one runs the barrier)
    for dt in (torch.float64, torch.float32):
        n, ns = 100000, 70001
        rows = [torch.rand(2 * n, dtype=dt, device=cuda) for _ in range(5)]
        want = [r.clone() for r in rows]
        src = torch.randperm(n, device=cuda)[:ns].to(torch.int32)
        dst = (n + torch.arange(ns, device=cuda)).to(torch.int32)
        for k in range(5):
            want[k][n:n + ns] = rows[k][src.long()]
        tables = tb.RankTables([rows], cuda)
        v = torch.tensor([1.5], dtype=dt, device=cuda)
        out = torch.zeros(1, dtype=dt, device=cuda)
        for rep in range(3):                               # the counter is left at zero by every call
            mb.push_barrier(src, torch.zeros(ns, dtype=torch.int32, device=cuda), dst, rows, tables, v, out)
            mb.push_barrier(src, torch.zeros(ns, dtype=torch.int32, device=cuda), dst, rows, tables)
        torch.cuda.synchronize()
        assert float(out[0]) == 1.5 and int(mb.push_counter[0]) == 0
        for k in range(5):
            assert torch.equal(rows[k], want[k])
        empty = torch.zeros(0, dtype=torch.int32, device=cuda)
        mb.push_barrier(empty, empty, empty, rows, tables)  # nothing to send: still a barrier
        torch.cuda.synchronize()
    L = tb.lib()
    assert L.t8b200_peer_barrier(0, 0, C.c_longlong(1), None, None, 1, None, None) != 0
    assert L.t8b200_peer_barrier(2, 2, C.c_longlong(1), C.c_void_p(mb.table.data_ptr()), None, 1, None, None) != 0
    assert L.t8b200_peer_barrier(1, 0, C.c_longlong(0), C.c_void_p(mb.table.data_ptr()), None, 1, None, None) != 0
    mb.close()
