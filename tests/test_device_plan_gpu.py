"""f-2: the tile plan built on the device (t8b200_plan_create_device) from device-resident connectivity arrays equals the
host builder's plan array by array (structured-chunk records, halo lists in thread order, owner ranks, partition-boundary
flags, ghost-tail redirection and pull lists), gives bitwise the same states, and reports meshes it does not cover."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from util import rel_linf

pytestmark = pytest.mark.gpu


def host_arrays(conn_host, dtype, ghost_tail):
    """Arrays 13-15, 17-18 of the HOST builder (host-only plan: nothing uploaded)."""
    import t8gpu_b200 as tb
    lib = tb.lib()
    npdt = np.float64 if dtype == torch.float64 else np.float32

    def arr(k, dt):
        v = conn_host.get(k)
        return None if v is None or len(v) == 0 else np.ascontiguousarray(v, dtype=dt)

    def p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    keep = [arr("face_neighbors", np.int32), arr("face_normals", npdt), arr("face_areas", npdt), arr("ranks", np.int32),
            arr("indices", np.int32), arr("x_face_neighbors", np.int32), arr("x_face_normals", npdt),
            arr("x_face_areas", npdt)]
    ng = int(conn_host.get("n_ghost", 0))
    h = C.c_void_p()
    fn = lib.t8b200_plan_create_ghost_tail_host if ghost_tail else lib.t8b200_plan_create_host
    assert fn(C.byref(h), int(dtype == torch.float64), C.c_int64(int(conn_host["n_local"])), C.c_int64(ng),
              int(conn_host["n_faces"]), int(conn_host["n_bfaces"]), p(keep[0]), p(keep[1]), p(keep[2]),
              p(keep[3]) if ng else None, p(keep[4]) if ng else None, int(conn_host.get("n_xfaces", 0)), p(keep[5]),
              p(keep[6]), p(keep[7])) == 0
    out = {}
    for which in (13, 14, 15, 17, 18):
        data, count, eb = C.c_void_p(), C.c_int64(), C.c_int()
        assert lib.t8b200_plan_host_array(h, which, C.byref(data), C.byref(count), C.byref(eb)) == 0
        n = count.value
        out[which] = (np.frombuffer((C.c_char * (n * 4)).from_address(data.value), dtype=np.int32).copy() if n
                      else np.zeros(0, np.int32))
    lib.t8b200_plan_destroy(h)
    return out


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("level", [4, 5])
def test_uniform_forest_plan_equals_host_builder(cuda, level, dtype):
    import t8gpu_b200 as tb
    conn = tb.cartesian_uniform_connectivity(3, level, dtype, 1, 0, device=cuda)
    plan = tb.Plan.from_device(conn, dtype)
    assert plan is not None and plan.info["n_chunks"] == int(conn["n_local"]) // 256
    host = host_arrays(tb.conn_to_host(conn), dtype, False)
    for which in (13, 14):
        assert np.array_equal(plan.device_array(which), host[which]), which
    assert plan.device_array(15).size == 0 and plan.n_tail == 0
    # the same states, bit for bit, as with the host-built plan; criteria agree to rounding (different summation order)
    a = tb.EulerSolver(tb.conn_to_host(conn), conn["volumes"], dtype, device=cuda)
    b = tb.EulerSolver(None or tb.conn_to_host(conn), conn["volumes"], dtype, device=cuda, plan=plan)
    tb.init_kelvin_helmholtz(3, conn["centroids"], a.variables(a.next))
    rng = np.random.default_rng(3)
    u0 = a.state().cpu().numpy() * (1 + 0.02 * rng.uniform(-1, 1, (5, a.n)))
    a.set_state(u0)
    b.set_state(u0)
    dt = 0.1 * 2.0 ** -level
    for _ in range(3):
        a.iterate(dt)
        b.iterate(dt)
    assert torch.equal(a.state(), b.state()) and torch.equal(a.max_wave_speed(), b.max_wave_speed())
    ca = tb.gradient_criteria(a.plan, a.state()[0], a.volume()).cpu().numpy()
    cb = tb.gradient_criteria(plan, b.state()[0], b.volume()).cpu().numpy()
    assert np.abs(ca - cb).max() <= (1e-13 if dtype == torch.float64 else 1e-5) * np.abs(ca).max()


@pytest.mark.parametrize("ghost_tail", [False, True])
@pytest.mark.parametrize("brick,P", [((2, 1, 1), 2), ((2, 2, 1), 4)])
def test_partitioned_brick_plan_equals_host_builder(cuda, brick, P, ghost_tail):
    import t8gpu_b200 as tb
    dtype = torch.float64
    for r in range(P):
        conn = tb.cartesian_uniform_connectivity(3, 4, dtype, P, r, device=cuda, brick=brick)
        plan = tb.Plan.from_device(conn, dtype, ghost_tail=ghost_tail)
        assert plan is not None
        host = host_arrays(tb.conn_to_host(conn), dtype, ghost_tail)
        for which in (13, 14, 15, 17, 18):
            assert np.array_equal(plan.device_array(which), host[which]), (which, r)
        if ghost_tail:
            assert plan.n_tail == host[17].size == int(conn["n_ghost"])
            assert (plan.device_array(15) == r).all()
        else:
            assert plan.device_array(13).reshape(-1, 4)[:, 3].sum() > 0      # partition-boundary chunks are flagged


HOST_DT = {1: np.uint8, 2: np.uint16, 4: np.int32, 8: np.float64}


def host_all(conn_host, dtype, ghost_tail):
    """All 20 arrays of the HOST builder + plan info (floating-point arrays in the plan's precision)."""
    import t8gpu_b200 as tb
    lib = tb.lib()
    npdt = np.float64 if dtype == torch.float64 else np.float32

    def arr(k, dt):
        v = conn_host.get(k)
        return None if v is None or len(v) == 0 else np.ascontiguousarray(v, dtype=dt)

    def p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    keep = [arr("face_neighbors", np.int32), arr("face_normals", npdt), arr("face_areas", npdt), arr("ranks", np.int32),
            arr("indices", np.int32), arr("x_face_neighbors", np.int32), arr("x_face_normals", npdt),
            arr("x_face_areas", npdt)]
    ng = int(conn_host.get("n_ghost", 0))
    h = C.c_void_p()
    fn = lib.t8b200_plan_create_ghost_tail_host if ghost_tail else lib.t8b200_plan_create_host
    assert fn(C.byref(h), int(dtype == torch.float64), C.c_int64(int(conn_host["n_local"])), C.c_int64(ng),
              int(conn_host["n_faces"]), int(conn_host["n_bfaces"]), p(keep[0]), p(keep[1]), p(keep[2]),
              p(keep[3]) if ng else None, p(keep[4]) if ng else None, int(conn_host.get("n_xfaces", 0)), p(keep[5]),
              p(keep[6]), p(keep[7])) == 0
    out = {}
    for which in range(20):
        data, count, eb = C.c_void_p(), C.c_int64(), C.c_int()
        assert lib.t8b200_plan_host_array(h, which, C.byref(data), C.byref(count), C.byref(eb)) == 0
        n = count.value
        a = (np.frombuffer((C.c_char * (n * eb.value)).from_address(data.value), dtype=HOST_DT[eb.value]).copy() if n
             else np.zeros(0, HOST_DT[eb.value]))
        out[which] = a.astype(npdt) if which in (8, 9, 10, 11, 12) else a
    info = (C.c_int64 * 8)()
    assert lib.t8b200_plan_info(h, info) == 0
    out["info"] = [info[i] for i in (0, 1, 2, 3, 5, 6, 7)]          # without the device byte count
    lib.t8b200_plan_destroy(h)
    return out


def device_all(plan, dtype):
    import t8gpu_b200 as tb
    lib = tb.lib()
    lib.t8b200_plan_device_bytes.restype = C.c_int64
    npdt = np.float64 if dtype == torch.float64 else np.float32
    dts = {0: np.int32, 1: np.int32, 2: np.int32, 3: np.int32, 4: np.uint8, 5: np.uint16, 6: np.uint16, 7: np.uint16}
    out = {}
    for which in range(20):
        dt = npdt if which in (8, 9, 10, 11, 12) else dts.get(which, np.int32)
        n = lib.t8b200_plan_device_bytes(plan._h, which, None, C.c_int64(0))
        assert n >= 0, which
        a = np.zeros(n // np.dtype(dt).itemsize, dt)
        if n:
            assert lib.t8b200_plan_device_bytes(plan._h, which, a.ctypes.data_as(C.c_void_p), C.c_int64(n)) == n
        out[which] = a
    i = plan.info
    out["info"] = [i["n_chunks"], i["max_halo"], i["max_faces"], i["smem_bytes"], i["face_records"], i["halo_entries"],
                   i["chunk"]]
    return out


def assert_same_plan(conn_host, dtype, ghost_tail, cuda, tag=""):
    import t8gpu_b200 as tb
    H = host_all(conn_host, dtype, ghost_tail)
    plan = tb.Plan.from_device(tb.conn_to_device(conn_host, dtype, cuda), dtype, ghost_tail=ghost_tail)
    assert plan is not None, tag
    D = device_all(plan, dtype)
    assert D["info"] == H["info"], (tag, D["info"], H["info"])
    for which in range(20):
        assert D[which].size == H[which].size and np.array_equal(D[which].view(H[which].dtype), H[which]), (tag, which)
    return plan, H


def _amr_forest(level=3):
    f = oracle.Forest(3, level)
    for width in (0.2, 0.1):
        lv, cent, vol, _ = f.elements()
        f = f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < width, 20.0, 0.0), 10.0, 1, level + 2)
    return f


@pytest.mark.parametrize("mode", ["generic", "serial"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_generic_device_builder_equals_host_builder(cuda, dtype, mode, monkeypatch):
    """Hanging faces (overflow entries), walls, quads, boxes that are their own neighbours, general normals: the plan
    built by one warp per block (mode generic: device_plan.cu block_warp_kernel) or one thread per block (mode serial:
    the program of csrc/plan_block.cuh as the host emulation runs it) equals the host builder's in all 20 arrays."""
    from util import hybrid_mesh
    monkeypatch.setenv("T8B200_DEVICE_PLAN", mode)               # not the three-kernel builder of structured-only meshes
    npdt = np.float64 if dtype == torch.float64 else np.float32
    amr = _amr_forest()
    plan, H = assert_same_plan(amr.connectivity(dtype=npdt), dtype, False, cuda, "amr")
    assert H[7].size > 0                                          # overflow CSR in use
    assert_same_plan(oracle.Forest(3, 4, periodic=False).connectivity(dtype=npdt), dtype, False, cuda, "walls")
    assert_same_plan(oracle.Forest(3, 3).connectivity(dtype=npdt), dtype, False, cuda, "level 3")
    assert_same_plan(oracle.Forest(2, 5).connectivity(dtype=npdt), dtype, False, cuda, "quads")
    for per in (True, False):
        _, H = assert_same_plan(hybrid_mesh(12, per, npdt, seed=3, shuffle=not per)[0], dtype, False, cuda, "hybrid")
        assert H[9].size > 0                                      # general geometry arrays
    # the uniform forest through the generic builder: every chunk structured
    _, H = assert_same_plan(oracle.Forest(3, 4).connectivity(dtype=npdt), dtype, False, cuda, "uniform")
    assert H[13].size == 4 * 16 and H[16].size == 0
    monkeypatch.setenv("T8B200_TEST_MAX_HALO", "40")              # blocks split into 8+ chunks each
    _, H = assert_same_plan(oracle.Forest(3, 4).connectivity(dtype=npdt), dtype, False, cuda, "split")
    assert H["info"][0] > 16


def test_generic_device_builder_edge_cases(cuda, monkeypatch):
    """More than 256 distinct areas on axis-aligned faces (the compressed geometry does not apply), a mesh smaller than
    one block, a rank without elements."""
    import t8gpu_b200 as tb
    monkeypatch.setenv("T8B200_DEVICE_PLAN", "generic")
    conn = oracle.Forest(3, 4).connectivity(dtype=np.float64)
    rng = np.random.default_rng(5)
    conn["face_areas"] = conn["face_areas"] * (1.0 + rng.integers(0, 300, conn["face_areas"].size) / 1024.0)
    _, H = assert_same_plan(conn, torch.float64, False, cuda, "300 areas")
    assert H[9].size > 0 and H[4].size == 0                       # general-geometry arrays, no area indices
    conn = oracle.Forest(3, 4).connectivity(dtype=np.float64)
    conn["face_areas"] = conn["face_areas"] * (1.0 + rng.integers(0, 200, conn["face_areas"].size) / 1024.0)
    _, H = assert_same_plan(conn, torch.float64, False, cuda, "200 areas")
    assert H[8].size == 200 and H[4].size > 0                     # area table in use, no chunk is structured
    assert H[13].size == 0
    assert_same_plan(oracle.Forest(3, 2).connectivity(dtype=np.float32), torch.float32, False, cuda, "64 elements")
    assert_same_plan(oracle.Forest(2, 1).connectivity(dtype=np.float64), torch.float64, False, cuda, "4 quads")
    empty = dict(n_local=0, n_ghost=0, n_faces=0, n_bfaces=0, n_xfaces=0)
    assert tb.Plan.from_device(empty, torch.float64) is None      # cudaErrorNotSupported: the host builder's case


@pytest.mark.parametrize("ghost_tail", [False, True])
@pytest.mark.parametrize("P", [2, 3])
def test_generic_device_builder_multi_rank(cuda, P, ghost_tail, monkeypatch):
    """Partitioned adaptive forest: owner ranks, partition-boundary flags and launch order, ghost-tail redirection and
    pull lists as the host builder's."""
    monkeypatch.setenv("T8B200_DEVICE_PLAN", "generic")
    amr = _amr_forest()
    for dtype, npdt in ((torch.float64, np.float64), (torch.float32, np.float32)):
        for r in range(P):
            plan, H = assert_same_plan(amr.connectivity(P, r, dtype=npdt), dtype, ghost_tail, cuda, "P%d r%d" % (P, r))
            if ghost_tail:
                assert plan.n_tail == H[17].size > 0
    # a partition of a uniform forest whose boxes are not all structured (level 3) and one that is (level 4, ragged cut)
    for level in (3, 4):
        f = oracle.Forest(3, level)
        for r in range(P):
            assert_same_plan(f.connectivity(P, r, dtype=np.float64), torch.float64, ghost_tail, cuda, "uniform")


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_generic_device_plan_steps_like_the_host_plan(cuda, dtype):
    """leaves -> device connectivity -> device plan -> 3 steps: bitwise the states of the host-built plan."""
    import t8gpu_b200 as tb
    from util import perturbed_kh
    npdt = np.float64 if dtype == torch.float64 else np.float32
    f = _amr_forest()
    lv, cent, vol, _ = f.elements()
    conn = tb.forest_connectivity(3, True, tb.morton_keys(3, lv, cent), lv, dtype, device=cuda)
    plan = tb.Plan.from_device(conn, dtype)
    assert plan is not None and plan.info["built_on"] == "device"
    host = tb.conn_to_host(conn)
    a = tb.EulerSolver(host, conn["volumes"], dtype, device=cuda)
    b = tb.EulerSolver(host, conn["volumes"], dtype, device=cuda, plan=plan)
    u0, _ = perturbed_kh(f, npdt, seed=5)
    a.set_state(u0)
    b.set_state(u0)
    for _ in range(3):
        a.iterate(1e-3)
        b.iterate(1e-3)
    assert torch.isfinite(a.state()).all() and torch.equal(a.state(), b.state())
    ca = tb.gradient_criteria(a.plan, a.state()[0], a.volume())
    cb = tb.gradient_criteria(plan, b.state()[0], b.volume())
    assert torch.equal(ca, cb)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_split_stage_passes_equal_the_single_launch(cuda, dtype):
    """t8b200_fused_stage_part: interior pass + boundary pass of a ghost-tail plan write bitwise what the single launch
    writes (two emulated ranks of the brick (2,1,1), every stage, wave-speed maximum included); the boundary list of
    the device-built plan equals the host builder's; plans that are not structured-only report NotSupported."""
    import t8gpu_b200 as tb
    from t8gpu_b200.solver import NB_STEPS, NVAR
    P, L, brick = 2, 4, (2, 1, 1)
    dt = 0.1 * 2.0 ** -L
    conns = [tb.cartesian_uniform_connectivity(3, L, dtype, P, r, device=cuda, brick=brick) for r in range(P)]
    plans = [tb.Plan.from_device(c, dtype, ghost_tail=True) for c in conns]
    ns = [int(c["n_local"]) for c in conns]
    for r in range(P):
        host = host_arrays(tb.conn_to_host(conns[r]), dtype, True)
        bl = plans[r].device_array(19)
        flags = plans[r].device_array(13).reshape(-1, 4)[:, 3]
        assert np.array_equal(bl, np.nonzero(flags)[0]) and len(bl) > 0

    def run(split):
        bufs = []
        for r in range(P):
            b = torch.zeros((NVAR * NB_STEPS + 1, ns[r] + plans[r].n_tail), dtype=dtype, device=cuda)
            b[NVAR * NB_STEPS, :ns[r]] = conns[r]["volumes"]
            tb.init_kelvin_helmholtz(3, conns[r]["centroids"], [b[k, :ns[r]] for k in range(5)])
            bufs.append(b)
        tabs = {s: tb.RankTables([[bufs[r][s * NVAR + k] for k in range(NVAR)] for r in range(P)], cuda)
                for s in range(NB_STEPS)}
        vmax = [torch.zeros(1, dtype=dtype, device=cuda) for _ in range(P)]
        nxt, prv = 0, 3
        for it in range(3):
            nxt, prv = prv, nxt
            for stage, sin, sout in ((1, prv, 1), (2, 1, 2), (3, 2, nxt)):
                for r in range(P):
                    plans[r].pull([bufs[r][sin * NVAR + k] for k in range(NVAR)], tabs[sin])
                for r in range(P):
                    v = lambda s: [bufs[r][s * NVAR + k, :ns[r]] for k in range(NVAR)]  # noqa: E731
                    sm = vmax[r] if stage == 3 else None
                    if split:
                        if sm is not None:
                            sm.zero_()
                        assert plans[r].stage_part(stage, 2, v(sin), v(prv), v(sout), bufs[r][NVAR * NB_STEPS], dt, speed_max=sm)
                        assert plans[r].stage_part(stage, 1, v(sin), v(prv), v(sout), bufs[r][NVAR * NB_STEPS], dt, speed_max=sm)
                    else:
                        plans[r].stage(stage, v(sin), v(prv), v(sout), bufs[r][NVAR * NB_STEPS], dt, speed_max=sm)
        torch.cuda.synchronize()
        return [b[nxt * NVAR:(nxt + 1) * NVAR, :ns[r]].clone() for r, b in enumerate(bufs)], [float(x) for x in vmax]

    whole, vm0 = run(False)
    parts, vm1 = run(True)
    for r in range(P):
        assert torch.equal(whole[r], parts[r])
    assert vm0 == vm1 and min(vm0) > 0

    # t8b200_fused_stage_push: the push folded into the stage kernel (no pull / push kernel at all) == the pulled run
    from t8gpu_b200 import multi

    class OneProcess:        # send_lists gathers the pull lists of all ranks: here they are all in this process
        def __init__(self):
            self.calls = []

        def all_gather(self, out, mine):
            self.calls.append((out, mine))

    def lists_for(rank):
        src, drk, dix = [], [], []
        for p in range(P):
            if p == rank:
                continue
            rk, ix = plans[p].device_array(17), plans[p].device_array(18)
            sel = np.nonzero(rk == rank)[0]
            src.append(ix[sel]); drk.append(np.full(len(sel), p, np.int32)); dix.append((ns[p] + sel).astype(np.int32))
        cat = lambda a: torch.as_tensor(np.concatenate(a)).to(torch.int32).to(cuda)  # noqa: E731
        return cat(src), cat(drk), cat(dix)

    csr = [multi.send_csr(lists_for(r), ns[r], cuda) for r in range(P)]
    bufs = []
    for r in range(P):
        b = torch.zeros((NVAR * NB_STEPS + 1, ns[r] + plans[r].n_tail), dtype=dtype, device=cuda)
        b[NVAR * NB_STEPS, :ns[r]] = conns[r]["volumes"]
        tb.init_kelvin_helmholtz(3, conns[r]["centroids"], [b[k, :ns[r]] for k in range(5)])
        bufs.append(b)
    tabs = {s: tb.RankTables([[bufs[r][s * NVAR + k] for k in range(NVAR)] for r in range(P)], cuda)
            for s in range(NB_STEPS)}
    for r in range(P):      # the initial state reaches the peers' copies through the stand-alone push kernel
        tb.ghost_push(*lists_for(r), [bufs[r][k] for k in range(NVAR)], tabs[0])
    vmax = [torch.zeros(1, dtype=dtype, device=cuda) for _ in range(P)]
    nxt, prv = 0, 3
    for it in range(3):
        nxt, prv = prv, nxt
        for stage, sin, sout in ((1, prv, 1), (2, 1, 2), (3, 2, nxt)):
            for r in range(P):
                v = lambda s: [bufs[r][s * NVAR + k, :ns[r]] for k in range(NVAR)]  # noqa: E731
                assert plans[r].stage_push(stage, v(sin), v(prv), v(sout), tabs[sout], bufs[r][NVAR * NB_STEPS], dt,
                                           csr[r], speed_max=vmax[r] if stage == 3 else None)
    torch.cuda.synchronize()
    for r in range(P):
        assert torch.equal(bufs[r][nxt * NVAR:(nxt + 1) * NVAR, :ns[r]], whole[r])
    assert [float(x) for x in vmax] == vm0
    # a plan with generic chunks (level 3: a tree 8 elements across is its own neighbour) does not support the split
    c3 = tb.cartesian_uniform_connectivity(3, 3, dtype, P, 0, device=cuda, brick=brick)
    p3 = tb.Plan(tb.conn_to_host(c3), dtype, ghost_tail=True)
    b3 = torch.zeros((NVAR * NB_STEPS + 1, int(c3["n_local"]) + p3.n_tail), dtype=dtype, device=cuda)
    v3 = lambda s: [b3[s * NVAR + k, :int(c3["n_local"])] for k in range(NVAR)]  # noqa: E731
    assert p3.stage_part(1, 1, v3(0), None, v3(1), b3[NVAR * NB_STEPS], dt) is False


def subgrid_host_arrays(conn_host, volumes, dtype):
    import t8gpu_b200 as tb
    lib = tb.lib()
    npdt = np.float64 if dtype == torch.float64 else np.float32

    def arr(k, dt):
        v = conn_host.get(k)
        return None if v is None or len(v) == 0 else np.ascontiguousarray(v, dtype=dt)

    def p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    keep = [arr("face_neighbors", np.int32), arr("face_normals", npdt), arr("face_areas", npdt),
            arr("level_diff", np.int32), arr("offsets", np.int32), np.ascontiguousarray(volumes, dtype=npdt),
            arr("ranks", np.int32), arr("indices", np.int32), arr("x_face_neighbors", np.int32),
            arr("x_face_normals", npdt), arr("x_face_areas", npdt), arr("x_level_diff", np.int32),
            arr("x_offsets", np.int32)]
    ng = int(conn_host.get("n_ghost", 0))
    sh = C.c_void_p()
    assert lib.t8b200_subgrid_plan_create_host(
        C.byref(sh), int(dtype == torch.float64), 3, C.c_int64(int(conn_host["n_local"])), C.c_int64(ng),
        int(conn_host["n_faces"]), int(conn_host["n_bfaces"]), p(keep[0]), p(keep[1]), p(keep[2]), p(keep[3]), p(keep[4]),
        p(keep[5]), p(keep[6]) if ng else None, p(keep[7]) if ng else None, int(conn_host.get("n_xfaces", 0)), p(keep[8]),
        p(keep[9]), p(keep[10]), p(keep[11]), p(keep[12])) == 0
    lib.t8b200_subgrid_plan_base.restype = C.c_void_p
    base = C.c_void_p(lib.t8b200_subgrid_plan_base(sh))
    out = {}
    for which in (13, 14, 15):
        data, count, eb = C.c_void_p(), C.c_int64(), C.c_int()
        assert lib.t8b200_plan_host_array(base, which, C.byref(data), C.byref(count), C.byref(eb)) == 0
        n = count.value
        out[which] = (np.frombuffer((C.c_char * (n * 4)).from_address(data.value), dtype=np.int32).copy() if n
                      else np.zeros(0, np.int32))
    lib.t8b200_subgrid_plan_destroy(sh)
    return out


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("P", [1, 2])
def test_subgrid_cell_plan_equals_host_builder(cuda, P, dtype):
    """t8b200_subgrid_plan_create_device: Subgrid<4,4,4> cell-level plan from device arrays == the host builder's
    (records, halo lists in thread order, owner ranks), same states bit for bit; an adapted forest is reported."""
    import t8gpu_b200 as tb
    from bench_subgrid import subgrid_connectivity
    brick = (2, 1, 1) if P == 2 else (1, 1, 1)
    for r in range(P):
        conn = subgrid_connectivity(3, dtype, r, P, cuda, brick)
        plan = tb.SubgridPlan.from_device(conn, conn["volumes"], dtype)
        assert plan is not None and plan.info["n_chunks"] == int(conn["n_local"]) // 4
        host = subgrid_host_arrays(tb.conn_to_host(conn), conn["volumes"].cpu().numpy(), dtype)
        for which in (13, 14, 15):
            assert np.array_equal(plan.device_array(which), host[which]), (which, r)
    if P == 1:
        f = oracle.Forest(3, 3)
        npdt = np.float64 if dtype == torch.float64 else np.float32
        lv, cent, vol, _ = f.elements()
        a = tb.SubgridEulerSolver(f.connectivity(subgrid=True, dtype=npdt), vol.astype(npdt), dtype, device=cuda, mode="fused")
        b = tb.SubgridEulerSolver(f.connectivity(subgrid=True, dtype=npdt), vol.astype(npdt), dtype, device=cuda, mode="fused")
        b.plan = plan
        u0 = oracle.subgrid_init_kh(3, cent.astype(npdt), lv, npdt)
        a.set_state(u0)
        b.set_state(u0)
        for _ in range(3):
            a.iterate(1e-3)
            b.iterate(1e-3)
        assert torch.equal(a.state(), b.state())


def subgrid_host_all(conn_host, volumes, dtype, dim, ghost_tail):
    """All 20 arrays + info of the HOST builder's cell-level plan."""
    import t8gpu_b200 as tb
    lib = tb.lib()
    npdt = np.float64 if dtype == torch.float64 else np.float32

    def arr(k, dt):
        v = conn_host.get(k)
        return None if v is None or len(v) == 0 else np.ascontiguousarray(v, dtype=dt)

    def p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    keep = [arr("face_neighbors", np.int32), arr("face_normals", npdt), arr("face_areas", npdt),
            arr("level_diff", np.int32), arr("offsets", np.int32), np.ascontiguousarray(volumes, dtype=npdt),
            arr("ranks", np.int32), arr("indices", np.int32), arr("x_face_neighbors", np.int32),
            arr("x_face_normals", npdt), arr("x_face_areas", npdt), arr("x_level_diff", np.int32),
            arr("x_offsets", np.int32)]
    ng = int(conn_host.get("n_ghost", 0))
    sh = C.c_void_p()
    fn = lib.t8b200_subgrid_plan_create_ghost_tail_host if ghost_tail else lib.t8b200_subgrid_plan_create_host
    assert fn(C.byref(sh), int(dtype == torch.float64), dim, C.c_int64(int(conn_host["n_local"])), C.c_int64(ng),
              int(conn_host["n_faces"]), int(conn_host["n_bfaces"]), p(keep[0]), p(keep[1]), p(keep[2]), p(keep[3]),
              p(keep[4]), p(keep[5]), p(keep[6]) if ng else None, p(keep[7]) if ng else None,
              int(conn_host.get("n_xfaces", 0)), p(keep[8]), p(keep[9]), p(keep[10]), p(keep[11]), p(keep[12])) == 0
    lib.t8b200_subgrid_plan_base.restype = C.c_void_p
    base = C.c_void_p(lib.t8b200_subgrid_plan_base(sh))
    out = {}
    for which in range(20):
        data, count, eb = C.c_void_p(), C.c_int64(), C.c_int()
        assert lib.t8b200_plan_host_array(base, which, C.byref(data), C.byref(count), C.byref(eb)) == 0
        n = count.value
        a = (np.frombuffer((C.c_char * (n * eb.value)).from_address(data.value), dtype=HOST_DT[eb.value]).copy() if n
             else np.zeros(0, HOST_DT[eb.value]))
        out[which] = a.astype(npdt) if which in (8, 9, 10, 11, 12) else a
    info = (C.c_int64 * 8)()
    assert lib.t8b200_plan_info(base, info) == 0
    out["info"] = [info[i] for i in (0, 1, 2, 3, 5, 6, 7)]
    lib.t8b200_subgrid_plan_destroy(sh)
    return out


class _Base:   # device_all() reads plan._h: the cell-level plan behind a subgrid plan
    def __init__(self, plan):
        self._h, self.info = plan._base, plan.info


@pytest.mark.parametrize("ghost_tail", [False, True])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_generic_device_builder_on_subgrid_cells(cuda, dtype, ghost_tail, monkeypatch):
    """Cell-level plans of adapted / walled Subgrid<4,4,4> and Subgrid<4,4> forests built on the device (one warp per 256
    cells over the cell faces) equal the host builder's in all 20 arrays, one rank and partitioned."""
    import t8gpu_b200 as tb
    monkeypatch.setenv("T8B200_DEVICE_PLAN", "generic")
    npdt = np.float64 if dtype == torch.float64 else np.float32
    cases = []
    f = oracle.Forest(3, 2)
    cases.append((f, 3))
    lv, cent, vol, _ = f.elements()
    cases.append((f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < 0.3, 20.0, 0.0), 10.0, 1, 4), 3))
    g = oracle.Forest(3, 1, periodic=False)
    lv, cent, vol, _ = g.elements()
    cases.append((g.adapt(np.where(cent[:, 2] < 0.5, 1.0, 0.0), 0.02, 1, 2), 3))
    h = oracle.Forest(2, 2, periodic=False)
    lv, cent, vol, _ = h.elements()
    cases.append((h.adapt(np.where(cent[:, 1] < 0.5, 1.0, 0.0), 0.02, 1, 3), 2))
    for forest, dim in cases:
        vol = forest.elements()[2]
        for P in ((1, 2) if not ghost_tail else (2,)):
            off = forest.partition_offsets(P)
            for r in range(P):
                conn = forest.connectivity(P, r, subgrid=True, dtype=npdt)
                lvol = vol[off[r]:off[r + 1]].astype(npdt)
                H = subgrid_host_all(conn, lvol, dtype, dim, ghost_tail)
                plan = tb.SubgridPlan.from_device(tb.conn_to_device(conn, dtype, cuda), torch.as_tensor(lvol).to(cuda),
                                                  dtype, ghost_tail=ghost_tail)
                assert plan is not None
                D = device_all(_Base(plan), dtype)
                assert D["info"] == H["info"], (dim, P, r, D["info"], H["info"])
                for which in range(20):
                    assert D[which].size == H[which].size and np.array_equal(D[which].view(H[which].dtype), H[which]), \
                        (which, dim, P, r)
