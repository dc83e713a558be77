"""The N > 1 host logic under torch.distributed with the gloo backend, world_size 2, on CPU (no GPU, no CUDA extension
calls): wire exchange and pointer tables, consistency of the two ranks' partition connectivity (ghost tables, mirrored
partition-boundary faces), the owner-computes stage protocol restated with the oracle arithmetic (both ranks evaluate
their shared faces, results identical to the single-rank run), and the global wave-speed reduction."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, HERE)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import oracle
        from t8gpu_b200 import multi
        from util import perturbed_kh
        out = {}

        # ---- 1. wire exchange + pointer tables
        handle = bytes([(rank * 37 + i) % 256 for i in range(64)])
        cap = 1024 + 32 * rank
        wires = multi.exchange_wires(dist, handle, cap, world, "cpu")
        assert [c for _, c in wires] == [1024, 1056]
        assert wires[rank][0] == handle and wires[1 - rank][0] != handle
        rows = multi.row_pointers([1 << 20, 1 << 30], [c for _, c in wires], 5, 5, 8)
        assert rows[2][3][1] == (1 << 30) + (2 * 5 + 3) * 1056 * 8 and rows[0][0][0] == 1 << 20

        # ---- 2. the two partitions of one adapted forest
        f = oracle.Forest(3, 2)
        lv, cent, vol, _ = f.elements()
        f = f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < 0.3, 20.0, 0.0), 10.0, 1, 3, nranks=world)
        lv, cent, vol, _ = f.elements()
        off = f.partition_offsets(world)
        conn = f.connectivity(world, rank, dtype=np.float64)
        nl, ng = int(conn["n_local"]), int(conn["n_ghost"])
        assert nl == off[rank + 1] - off[rank]
        counts = [None] * world
        dist.all_gather_object(counts, nl)
        # every ghost resolves to a live element of the other rank, with the global id the forest says
        for g in range(ng):
            r, i = int(conn["ranks"][nl + g]), int(conn["indices"][nl + g])
            assert r != rank and 0 <= i < counts[r]
            assert off[r] + i == int(conn["ghost_global"][g])
        # partition-boundary faces: the lower rank lists them as regular faces, the higher rank as x-faces, same pairs
        def glob(e):
            return off[rank] + e if e < nl else int(conn["ghost_global"][e - nl])
        nf = int(conn["n_faces"])
        nbr = conn["face_neighbors"][:2 * nf].reshape(-1, 2)
        mine = sorted((min(glob(a), glob(b)), max(glob(a), glob(b))) for a, b in nbr if a >= nl or b >= nl)
        xn = conn["x_face_neighbors"].reshape(-1, 2)
        xs = sorted((min(glob(a), glob(b)), max(glob(a), glob(b))) for a, b in xn)
        both = [None] * world
        dist.all_gather_object(both, (mine, xs))
        assert both[0][0] == both[1][1] and len(both[0][0]) > 0      # rank 0 owns them, rank 1 mirrors them
        assert both[1][0] == [] and both[0][1] == []

        # ---- 3. owner-computes protocol with the oracle arithmetic: ghosts from the other rank, regular + x faces,
        # only local accumulators kept; one all-reduce per stage orders the ranks and carries the wave speed
        u0, volT = perturbed_kh(f, np.float64, seed=7)
        dt = 0.05 * 2.0 ** -3
        ref, _, _ = oracle.iterate(f.connectivity(dtype=np.float64), volT, u0, dt)
        xconn = dict(n_faces=int(conn["n_xfaces"]), n_bfaces=0, face_neighbors=conn["x_face_neighbors"],
                     face_normals=conn["x_face_normals"], face_areas=conn["x_face_areas"])
        v = np.ascontiguousarray(volT[off[rank]:off[rank + 1]])
        prev = np.ascontiguousarray(u0[:, off[rank]:off[rank + 1]])
        token = torch.zeros(1, dtype=torch.float64)

        def stage(k, cur, prev):
            every = [None] * world
            dist.all_gather_object(every, cur)                     # stands for the peer loads of the ghost states
            ext = np.zeros((5, nl + ng))
            ext[:, :nl] = cur
            for g in range(ng):
                ext[:, nl + g] = every[int(conn["ranks"][nl + g])][:, int(conn["indices"][nl + g])]
            flux = np.zeros_like(ext)
            sp = np.zeros(nf + int(conn["n_bfaces"]))
            oracle.flux_faces(conn, np.ascontiguousarray(ext), flux, sp)
            spx = np.zeros(max(1, xconn["n_faces"]))
            if xconn["n_faces"]:
                oracle.flux_faces(xconn, np.ascontiguousarray(ext), flux, spx)
            out_ = np.zeros_like(cur)
            fl = np.ascontiguousarray(flux[:, :nl])
            oracle.rk_stage(k, prev, cur, out_, fl, v, dt)
            multi.stage_barrier(dist, token)
            return out_, max(sp.max(initial=0.0), spx.max(initial=0.0))

        s1, _ = stage(1, prev, prev)
        s2, _ = stage(2, s1, prev)
        nxt, vmax = stage(3, s2, prev)
        assert np.abs(nxt - ref[:, off[rank]:off[rank + 1]]).max() <= 1e-13 * np.abs(ref).max()

        # ---- 4. global CFL reduction: both ranks end up with the same dt
        t = torch.tensor([vmax], dtype=torch.float64)
        multi.global_max_wave_speed(dist, t)
        allv = [None] * world
        dist.all_gather_object(allv, vmax)
        assert float(t[0]) == max(allv)
        out["dt"] = multi.timestep(float(t[0]), 0.7, 4)
        dts = [None] * world
        dist.all_gather_object(dts, out["dt"])
        assert dts[0] == dts[1] > 0

        # ---- 5. push lists = the peers' pull lists regrouped by owner (send_lists over an all-gather)
        class FakePlan:      # arrays 17 / 18 of a ghost-tail plan: sorted distinct (owner, index) pairs of the ghosts
            def __init__(self, ranks, idx):
                self.a = {17: np.asarray(ranks, np.int32), 18: np.asarray(idx, np.int32)}

            def device_array(self, which):
                return self.a[which]
        gh = sorted(set((int(conn["ranks"][nl + g]), int(conn["indices"][nl + g])) for g in range(ng)))
        plan = FakePlan([r for r, _ in gh], [i for _, i in gh])
        src, drk, dix = multi.send_lists(dist, plan, nl, rank, world, "cpu")
        lists = [None] * world
        dist.all_gather_object(lists, (gh, nl, src.tolist(), drk.tolist(), dix.tolist()))
        other = 1 - rank
        o_gh, o_nl = lists[other][0], lists[other][1]
        want = [(i, other, o_nl + j) for j, (r, i) in enumerate(o_gh) if r == rank]     # what the other rank pulls from me
        assert list(zip(src.tolist(), drk.tolist(), dix.tolist())) == want and len(want) > 0

        # ---- 6. one adapt + repartition cycle: the per-rank index arithmetic (adapt_partition_ranges) drives the
        # restated remaps (oracle.adapt_remap / partition_remap) to exactly the one-rank result
        lv, cent, vol, _ = f.elements()
        crit = np.where(np.abs(cent[:, 0] - 0.5) < 0.2, 20.0, 0.0)
        f2 = f.adapt(crit, 10.0, 1, 4, nranks=world)
        amap = f.adapt_map(f2)
        off2 = f2.partition_offsets(world)
        u_glob, vol_glob = oracle.adapt_remap(amap, u0, volT, 0)                  # one rank: the whole forest at once
        lo, ad, owner, index = multi.adapt_partition_ranges(amap, off, off2, rank)
        assert lo[0] == 0 and lo[-1] == f2.num_elements and ad[0] == 0 and ad[-1] == off[rank + 1] - off[rank]
        u_mid, vol_mid = oracle.adapt_remap(ad, np.ascontiguousarray(u0[:, off[rank]:off[rank + 1]]),
                                            np.ascontiguousarray(volT[off[rank]:off[rank + 1]]), 0)
        mids = [None] * world
        dist.all_gather_object(mids, (u_mid, vol_mid))                            # stands for the peer tables
        u_new, vol_new = oracle.partition_remap(owner, index, [m[0] for m in mids], [m[1] for m in mids])
        assert np.array_equal(u_new, u_glob[:, off2[rank]:off2[rank + 1]])
        assert np.array_equal(vol_new, vol_glob[off2[rank]:off2[rank + 1]])
        assert np.array_equal(vol_new, f2.elements()[2][off2[rank]:off2[rank + 1]])
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:   # pragma: no cover
        import traceback
        q.put((rank, "FAILED: " + "".join(traceback.format_exception(type(e), e, e.__traceback__))))


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def test_wire_format():
    sys.path.insert(0, ROOT)
    from t8gpu_b200 import multi
    w = multi.pack_wire(bytes(range(64)), 123456789012)
    assert len(w) == multi.WIRE_BYTES
    h, c = multi.unpack_wire(w)
    assert h == bytes(range(64)) and c == 123456789012
    with pytest.raises(ValueError):
        multi.pack_wire(b"short", 1)
    with pytest.raises(ValueError):
        multi.unpack_wire([0] * 10)
    assert multi.timestep(2.0, 0.7, 4) == 0.7 * 0.5 ** 4 / 2.0 and multi.timestep(2.0, 0.7, 4, dt_cap=1e-3) == 1e-3
