"""`bench.py --workload hybrid`: BASELINE config 5 -- compressible Euler on a MIXED-element mesh (hexahedra + prisms +
tetrahedra, general unit normals), STRONG scaling over 1..8 GPUs.

t8code's hybrid cmeshes are not available in this image, so the mesh is a flat-array one in the reference's
MeshConnectivityAccessor layout (t8gpu_b200/meshes.py): a conforming periodic hex / prism / tet pattern tiled to the
requested size, split into contiguous element ranges (one per GPU) with ghosts, lower-rank-owns faces and the x-faces of
the higher rank.  The kernels are element-type agnostic: this is the general-normal path of the tile plan (four geometry
values per face record, 130-op flux).  Ghosts: local tail filled by one pull kernel per stage over the peer tables.
Parity: the same run on a small tiling against the CPU oracle (the checker), tolerance per step from the north star.
End-to-end parity on real t8code mixed meshes is blocked on t8code."""
import json
import os
import time

import numpy as np
import torch


def pad32(n):
    return max(32, (n + 31) // 32 * 32)


class HybridRank:
    def __init__(self, tiles, dtype, rank, world, device, dist, pattern=6):
        import t8gpu_b200 as tb
        from t8gpu_b200 import meshes
        from t8gpu_b200.multi import exchange_wires
        self.tb, self.rank, self.world, self.device, self.dtype, self.dist = tb, rank, world, device, dtype, dist
        npdt = np.float64 if dtype == torch.float64 else np.float32
        esz = 8 if dtype == torch.float64 else 4
        t0 = time.time()
        c0, v0, x0, shift = meshes.hybrid_mesh(pattern, True, npdt, with_shift=True)
        big, vol, cent = meshes.tile_periodic_mesh(c0, v0, x0, shift, tiles)
        self.n_total, self.n_faces_total = int(big["n_local"]), int(big["n_faces"])
        self.global_conn, self.global_vol, self.global_cent = big, vol, cent
        conn, lvol = meshes.partition_flat_mesh(big, vol, world, rank)
        self.off = conn["offsets_global"]
        self.t_mesh = time.time() - t0
        self.n = int(conn["n_local"])
        t0 = time.time()
        # the mesh generator is host code (stand-in for a t8code cmesh): upload its arrays once, plan on the device
        self.plan = tb.Plan.from_device(tb.conn_to_device(conn, dtype, device), dtype, ghost_tail=world > 1)
        if self.plan is None:
            self.plan = tb.Plan(conn, dtype, ghost_tail=world > 1)
        torch.cuda.synchronize()
        self.t_plan = time.time() - t0
        self.conn = conn
        self.cap = pad32(self.n + self.plan.n_tail)
        self.shared, self.mail = None, None
        if world > 1:
            self.shared = tb.SharedBuffer(26 * self.cap * esz, device)
            self.buffer = self.shared.tensor((26, self.cap), dtype)
            wires = exchange_wires(dist, self.shared.handle, self.cap, world, device)
            caps = [c for _, c in wires]
            bases = [self.shared.ptr if r == rank else self.shared.open_peer(wires[r][0]) for r in range(world)]
            self.tables = {s: tb.PointerTables([[bases[r] + (s * 5 + k) * caps[r] * esz for r in range(world)]
                                                for k in range(5)], device) for s in range(4)}
            self.mail = tb.PeerMailboxes(rank, world, device)
            self.mail.exchange([h for h, _ in exchange_wires(dist, self.mail.handle, 0, world, device)])
        else:
            self.buffer = torch.zeros((26, self.cap), dtype=dtype, device=device)
            self.tables = {s: None for s in range(4)}
        self.buffer[25, :self.n] = torch.as_tensor(np.ascontiguousarray(lvol.astype(npdt))).to(device)
        u0 = meshes.smooth_state(cent[self.off[rank]:self.off[rank + 1]], npdt, seed=31, amp=0.0)
        self.buffer[0:5, :self.n] = torch.as_tensor(u0).to(device)
        self.next, self.prev = 0, 3
        self.speed_loc = torch.zeros(1, dtype=dtype, device=device)
        self.speed_max = torch.zeros(1, dtype=dtype, device=device)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            self.mail.barrier(self.speed_loc, self.speed_max)

    def vars(self, step):
        return [self.buffer[step * 5 + k, :self.n] for k in range(5)]

    def state(self):
        return self.buffer[self.next * 5:(self.next + 1) * 5, :self.n]

    def iterate(self, dt):
        self.next, self.prev = self.prev, self.next
        prev, s1, s2, nxt = (self.vars(s) for s in (self.prev, 1, 2, self.next))
        vol, P, mail, launches = self.buffer[25, :self.n], self.plan, self.mail, 0
        for stage, sin, vin, vout in ((1, self.prev, prev, s1), (2, 1, s1, s2), (3, 2, s2, nxt)):
            if mail is not None:
                P.pull(vin, self.tables[sin])
                launches += 1
            P.stage(stage, vin, prev if stage > 1 else None, vout, vol, dt,
                    speed_max=self.speed_loc if stage == 3 else None)
            launches += 1
            if mail is not None:
                if stage == 3:
                    mail.barrier(self.speed_loc, self.speed_max)
                else:
                    mail.barrier()
                launches += 1
        return launches

    def close(self):
        torch.cuda.synchronize()
        self.plan = None
        if self.world > 1:
            self.dist.barrier()
            self.mail.close()
            self.shared.close()


def parity_small(dtype_name, rank, world, device, dist, steps=3):
    """The N-rank product run on a small tiling (2 x 2 x 2 pattern blocks, 5 184 elements) against the CPU oracle."""
    import oracle                       # the checker
    dtype = torch.float64 if dtype_name == "f64" else torch.float32
    npdt = np.float64 if dtype_name == "f64" else np.float32
    tol = 1e-12 if dtype_name == "f64" else 1e-5
    sol = HybridRank(2, dtype, rank, world, device, dist)
    from t8gpu_b200 import meshes
    u = meshes.smooth_state(sol.global_cent, npdt, seed=31, amp=0.0)
    dt = 0.02 / 12
    errs = []
    for _ in range(steps):
        u, _, _ = oracle.iterate(sol.global_conn, sol.global_vol, u, dt)
        sol.iterate(dt)
        a = sol.state().cpu().numpy().astype(np.float64)
        b = u[:, sol.off[rank]:sol.off[rank + 1]].astype(np.float64)
        scale = np.abs(u).max(axis=1).astype(np.float64)
        scale = np.where(scale < 1e-3 * scale.max(), scale.max(), scale)
        errs.append(float((np.abs(a - b).max(axis=1) / scale).max()))
    sol.close()
    worst = torch.tensor(errs, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    errs = [float(x) for x in worst.cpu()]
    return {"vs": "CPU oracle (restatement of examples/compressible_euler/kernels.cu) on the same mixed mesh, %d elements"
                  % (648 * 8), "steps": steps, "rel_linf_after_step": errs, "tolerance_per_step": tol,
            "ok": all(e <= (k + 1) * tol for k, e in enumerate(errs)), "ranks": world}


def run_hybrid(args, rank, world, device, dist=None, emit=True):
    from bench import ClockSampler, measured_peak
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    tiles = args.level if args.level is not None else 16      # 648 * tiles^3 elements in total
    parity = None if getattr(args, "no_parity", False) else parity_small(args.dtype, rank, world, device, dist)
    sol = HybridRank(tiles, dtype, rank, world, device, dist)
    dt = 0.02 / (6 * tiles)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        sol.iterate(dt)
    barrier()
    sampler = ClockSampler(device.index or 0)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, tw0 = 0, time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        launches += sol.iterate(dt)
    ev1.record(stream)
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop(tw0, time.time()) if rank == 0 else None
    assert bool(torch.isfinite(sol.state()).all()), "the run diverged"
    esz = 8 if args.dtype == "f64" else 4
    n_tot, f_tot = sol.n_total, sol.n_faces_total
    # algorithmic bytes per element and step: state 8 V F + per stage volume F + faces/element x (8 + 4 F)
    alg = 40 * esz + 3 * (esz + f_tot / n_tot * (8 + 4 * esz))
    peak, src = measured_peak()
    ms_per_step = ms / args.steps
    achieved = alg * n_tot / world / (ms_per_step * 1e-3) / 1e9       # per GPU
    line = {"metric": "cell-updates/s per RK3 step", "value": n_tot * args.steps / (ms * 1e-3), "unit": "cell-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "compressible Euler on a mixed-element mesh (hexahedra + prisms + tetrahedra, general "
                                   "normals), %d elements, %d faces in total, %d GPU(s), fixed dt" % (n_tot, f_tot, world),
                       "mesh": "flat arrays in the MeshConnectivityAccessor layout: conforming periodic hex/prism/tet "
                               "pattern (648 elements) tiled %d^3 times; contiguous element ranges per GPU, ghosts + "
                               "x-faces; t8code hybrid cmeshes absent -> end-to-end parity blocked on t8code" % tiles,
                       "elements_per_gpu": sol.n, "ghosts_per_gpu": int(sol.conn["n_ghost"]),
                       "faces_per_element": f_tot / n_tot, "alg_bytes_per_element_step": alg,
                       "host_setup_s": {"mesh+partition_numpy": round(sol.t_mesh, 2), "tile_plan (" + sol.plan.info.get("built_on", "host") + ", incl. upload)": round(sol.t_plan, 3)},
                       "plan": sol.plan.info, "l2": "state %.0f MB per stage per GPU" % (2 * 5 * sol.n * esz / 1e6)},
            "clocks": clocks, "gpu_launches": launches, "parity": parity,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": src, "kernel": "fused_stage_kernel (general normals)",
                         "note": "per GPU"}}
    sol.close()
    if rank == 0 and emit:
        print(json.dumps(line))
    return line
