"""`bench.py --workload amr`: BASELINE config 3 -- Kelvin-Helmholtz on an ADAPTIVE hex forest, adapt + repartition every
`--every` steps, one t8code partition per GPU (1..8 GPUs, torchrun), ghost faces over NVLink.

Per cycle, on every rank:
  stepping      `every` RK3 steps of the fused path (ghost tail: barrier + pull + single-rank stage kernel per stage)
  criteria      t8b200_gradient_criteria on the tile plan (device), gathered over the ranks
  forest        adapt + 2:1 balance + repartition -- HOST work that stays with t8code (north star).  t8code is not
                installed in this image: oracle.Forest (the mini-forest restatement of the t8code semantics the reference
                relies on) stands in for it, replicated on every rank.  It is mesh management only; the
                stepping, the criteria, both remaps and the plan build timed here are the product.
  adapt remap   t8b200_adapt_remap: old local elements -> the adapted, not yet repartitioned elements (device)
  partition     t8b200_partition_remap: every rank PULLS its new elements from the ranks that hold them, through the
                peer tables over NVLink (mesh_manager.inl:625-643 does it through MPI + CUDA-IPC on one GPU)
  connectivity  the reference-layout arrays of the new partition: built on the device from the leaf list
                (t8b200_forest_connectivity; the subgrid variant still takes the host loop) + tile plan rebuild (device
                builder when every chunk is structured, host builder otherwise)
`--check`: every rank also steps the whole forest alone on its GPU and compares its partition after every cycle (the
N-rank run must agree with the one-rank run of the same forest sequence within the north-star tolerance per step).
"""
import json
import os
import time

import numpy as np
import torch

from t8gpu_b200.multi import adapt_partition_ranges

CRIT_SCALE, THRESHOLD = 10.0 / 0.5, 10.0   # the example's threshold (mesh_manager.inl:141) with the criterion scaled so
                                           # that the Cartesian shear layers refine (as tests/perf_amr.py)


def pad32(n):
    return max(32, (n + 31) // 32 * 32)


class RankMesh:
    """One rank's partition of one forest: connectivity, plan (ghost tail), MemoryManager-layout rows, peer tables."""

    def __init__(self, forest, rank, world, device, dtype, dist, rows=26, subgrid=False):
        """subgrid: Subgrid<4,4,4> cells (64 per element) in the variable rows, cell-level plan; the per-element volumes
        sit in the last row (first n entries) so that the partition remap can pull them from the peers."""
        import t8gpu_b200 as tb
        from t8gpu_b200.multi import exchange_wires
        self.tb, self.rank, self.world, self.device, self.dtype = tb, rank, world, device, dtype
        npdt = np.float64 if dtype == torch.float64 else np.float32
        esz = 8 if dtype == torch.float64 else 4
        t0 = time.time()
        self.off = forest.partition_offsets(world)
        lv, cent = forest.elements()[:2]
        self.S = 64 if subgrid else 1
        self.plan, conn_dev = None, None
        if subgrid:
            # element-level connectivity of the subgrid manager (level differences, neighbour offsets) on the device
            # from the leaf list (t8b200_forest_subgrid_connectivity), the cell-level plan on the device from it
            conn_dev = tb.forest_connectivity(3, True, tb.morton_keys(3, lv, cent), lv, dtype, world, rank, device=device,
                                              subgrid=True)
            torch.cuda.synchronize()
            self.t_conn, self.conn_on = time.time() - t0, "device"
            self.n = int(conn_dev["n_local"])
            t0 = time.time()
            self.plan = tb.SubgridPlan.from_device(conn_dev, conn_dev["volumes"], dtype, ghost_tail=world > 1)
            if self.plan is None:                         # a rank without elements
                self.plan = tb.SubgridPlan(tb.conn_to_host(conn_dev), conn_dev["volumes"].cpu().numpy(), dtype,
                                           ghost_tail=world > 1)
        else:
            # MeshManager connectivity on the device from the leaf list (t8b200_forest_connectivity): no host face loop
            conn_dev = tb.forest_connectivity(3, True, tb.morton_keys(3, lv, cent), lv, dtype, world, rank, device=device)
            torch.cuda.synchronize()
            self.t_conn, self.conn_on = time.time() - t0, "device"
            self.n = int(conn_dev["n_local"])
            t0 = time.time()
            # tile plan on the device as well (structured-only meshes: three kernels; else one thread per block)
            self.plan = tb.Plan.from_device(conn_dev, dtype, ghost_tail=world > 1)
            if self.plan is None:                         # a rank without elements
                self.plan = tb.Plan(tb.conn_to_host(conn_dev), dtype, ghost_tail=world > 1)
        self.nc = self.n * self.S                          # entries of a variable row (cells)
        torch.cuda.synchronize()
        self.t_plan = time.time() - t0
        self.cap = pad32(self.nc + self.plan.n_tail)
        self.rows = rows
        self.shared = None
        if world > 1:
            self.shared = tb.SharedBuffer(rows * self.cap * esz, device)
            self.buffer = self.shared.tensor((rows, self.cap), dtype)
            wires = exchange_wires(dist, self.shared.handle, self.cap, world, device)
            self.caps = [c for _, c in wires]
            self.bases = [self.shared.ptr if r == rank else self.shared.open_peer(wires[r][0]) for r in range(world)]
        else:
            self.buffer = torch.zeros((rows, self.cap), dtype=dtype, device=device)
            self.caps, self.bases = [self.cap], [self.buffer.data_ptr()]
        self._tables = {}
        self.esz = esz

    def vars(self, step):
        return [self.buffer[step * 5 + k, :self.nc] for k in range(5)]

    def volume(self):
        return self.buffer[25, :self.n] if self.rows == 26 else self.buffer[5, :self.n]

    def tables(self, step):
        """[var][rank] -> row address (MemoryAccessorAll of this step)."""
        if step not in self._tables:
            ptrs = [[self.bases[r] + (step * 5 + k) * self.caps[r] * self.esz for r in range(self.world)] for k in range(5)]
            self._tables[step] = self.tb.PointerTables(ptrs, self.device)
        return self._tables[step]

    def volume_table(self):
        vrow = 25 if self.rows == 26 else 5
        return torch.tensor([self.bases[r] + vrow * self.caps[r] * self.esz for r in range(self.world)],
                            dtype=torch.int64).to(self.device)

    def close(self):
        self.plan = None
        self._tables = {}
        if self.shared is not None:
            self.shared.close()
            self.shared = None


class AmrRun:
    def __init__(self, level, max_level, dtype, rank, world, device, dist, subgrid=False):
        import oracle                      # the t8code stand-in (host forest); see the module docstring
        import t8gpu_b200 as tb
        self.tb, self.oracle, self.dist, self.subgrid = tb, oracle, dist, subgrid
        self.rank, self.world, self.device, self.dtype, self.max_level = rank, world, device, dtype, max_level
        self.npdt = np.float64 if dtype == torch.float64 else np.float32
        self.forest = oracle.Forest(3, level)
        self.mesh = RankMesh(self.forest, rank, world, device, dtype, dist, subgrid=subgrid)
        lv, cent, vol, _ = self.forest.elements()
        o0, o1 = self.mesh.off[rank], self.mesh.off[rank + 1]
        self.mesh.volume().copy_(torch.as_tensor(vol[o0:o1].astype(self.npdt)))
        centres = torch.as_tensor(np.ascontiguousarray(cent[o0:o1].astype(self.npdt))).to(device).reshape(-1)
        if subgrid:                        # centres of the 4 x 4 x 4 cells (examples/subgrid/solver.inl:13-35)
            from bench_subgrid import cell_centers
            centres = cell_centers(centres, level, dtype)
        tb.init_kelvin_helmholtz(3, centres, self.mesh.vars(0))
        self.next, self.prev = 0, 3
        self.mail = None
        if world > 1:
            from t8gpu_b200.multi import exchange_wires
            self.mail = tb.PeerMailboxes(rank, world, device)
            self.mail.exchange([h for h, _ in exchange_wires(dist, self.mail.handle, 0, world, device)])
        self.speed_loc = torch.zeros(1, dtype=dtype, device=device)
        self.speed_max = torch.zeros(1, dtype=dtype, device=device)
        self.sync_all()

    def sync_all(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.mail.barrier(self.speed_loc, self.speed_max)

    def iterate(self, dt):
        m, P, mail = self.mesh, self.mesh.plan, self.mail
        self.next, self.prev = self.prev, self.next
        prev, s1, s2, nxt = (m.vars(s) for s in (self.prev, 1, 2, self.next))
        vol = m.volume()
        for stage, sin, vin, vout in ((1, self.prev, prev, s1), (2, 1, s1, s2), (3, 2, s2, nxt)):
            if mail is not None:
                P.pull(vin, m.tables(sin))
            if self.subgrid:
                P.stage(stage, vin, prev if stage > 1 else None, vout, vol, dt)
            else:
                P.stage(stage, vin, prev if stage > 1 else None, vout, vol, dt,
                        speed_max=self.speed_loc if stage == 3 else None)
            if mail is not None:
                if stage == 3:
                    mail.barrier(self.speed_loc, self.speed_max)
                else:
                    mail.barrier()

    def state(self):
        return self.mesh.buffer[self.next * 5:(self.next + 1) * 5, :self.mesh.nc]

    def adapt(self, T):
        """One adapt + repartition cycle; T accumulates the seconds of its parts."""
        tb, dist, world, rank, dev = self.tb, self.dist, self.world, self.rank, self.device
        m, f = self.mesh, self.forest
        # ---- criteria (device), gathered over the ranks
        t = time.time()
        if self.subgrid:      # H1 seminorm of the density inside each element (no ghosts), threshold of the example
            crit = tb.subgrid_criteria(3, m.buffer[self.next * 5, :m.nc], m.volume())
            threshold = 0.02
        else:
            if self.mail is not None:
                m.plan.pull(m.vars(self.next), m.tables(self.next))     # the criterion reads the ghosts' densities
            crit = tb.gradient_criteria(m.plan, m.buffer[self.next * 5, :m.n], m.volume()) * CRIT_SCALE
            threshold = THRESHOLD
        counts = np.diff(m.off)
        if world > 1:
            pad = torch.zeros(int(counts.max()), dtype=self.dtype, device=dev)
            pad[:m.n] = crit
            every = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(every, pad)
            crit_h = np.concatenate([every[q][:int(counts[q])].cpu().numpy() for q in range(world)])
        else:
            crit_h = crit.cpu().numpy()
        torch.cuda.synchronize()
        T["criteria_device+gather"] += time.time() - t
        # ---- forest: adapt + balance + repartition on the host (t8code's part; mini-forest stand-in)
        t = time.time()
        f2 = f.adapt(crit_h, threshold, 1, self.max_level, nranks=world)
        amap = f.adapt_map(f2)
        n_new = f2.num_elements
        off2 = f2.partition_offsets(world)
        lo, ad_h, owner, index = adapt_partition_ranges(amap, m.off, off2, rank)
        T["forest_host(t8code stand-in)"] += time.time() - t
        # ---- adapt remap into a peer-visible intermediate (5 variables + volume)
        t = time.time()
        n_mid = int(lo[rank + 1] - lo[rank])
        mid = RankMidBuffer(n_mid, rank, world, dev, self.dtype, dist, tb, m.S)
        ad = torch.as_tensor(ad_h).to(dev)
        tb.adapt_remap(ad, m.vars(self.next), mid.vars(), m.volume(), mid.volume(), 3 if self.subgrid else 0)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        T["adapt_remap_device"] += time.time() - t
        # ---- new partition: connectivity (host) + plan + rows
        new = RankMesh(f2, rank, world, dev, self.dtype, dist, subgrid=self.subgrid)
        T["connectivity (%s)" % new.conn_on] = T.get("connectivity (%s)" % new.conn_on, 0.0) + new.t_conn
        T["tile_plan"] += new.t_plan
        # ---- partition remap: pull the new elements from the ranks that hold them
        t = time.time()
        tb.partition_remap(torch.as_tensor(owner).to(dev), torch.as_tensor(index).to(dev), new.vars(0), mid.tables(),
                           new.volume(), mid.volume_table(), m.S)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        T["partition_remap_device"] += time.time() - t
        moved = int((owner != rank).sum())
        mid.close()
        m.close()
        self.mesh, self.forest = new, f2
        self.next, self.prev = 0, 3
        self.sync_all()
        return dict(elements=int(n_new), cells=int(n_new) * m.S, local=int(new.n), moved_between_ranks=moved,
                    chunks=int(new.plan.info["n_chunks"]), plan_built_on=new.plan.info.get("built_on", "host")), crit_h


class RankMidBuffer:
    """The adapted, not yet repartitioned elements of one rank: 5 variable rows + the volume row, peer-visible."""

    def __init__(self, n, rank, world, device, dtype, dist, tb, S=1):
        from t8gpu_b200.multi import exchange_wires
        self.n, self.world, self.device, self.tb, self.S = n, world, device, tb, S
        self.esz = 8 if dtype == torch.float64 else 4
        self.cap = pad32(n * S)
        self.shared = None
        if world > 1:
            self.shared = tb.SharedBuffer(6 * self.cap * self.esz, device)
            self.buffer = self.shared.tensor((6, self.cap), dtype)
            wires = exchange_wires(dist, self.shared.handle, self.cap, world, device)
            self.caps = [c for _, c in wires]
            self.bases = [self.shared.ptr if r == rank else self.shared.open_peer(wires[r][0]) for r in range(world)]
        else:
            self.buffer = torch.zeros((6, self.cap), dtype=dtype, device=device)
            self.caps, self.bases = [self.cap], [self.buffer.data_ptr()]

    def vars(self):
        return [self.buffer[k, :self.n * self.S] for k in range(5)]

    def volume(self):
        return self.buffer[5, :self.n]

    def tables(self):
        return self.tb.PointerTables([[self.bases[r] + k * self.caps[r] * self.esz for r in range(self.world)]
                                      for k in range(5)], self.device)

    def volume_table(self):
        return torch.tensor([self.bases[r] + 5 * self.caps[r] * self.esz for r in range(self.world)],
                            dtype=torch.int64).to(self.device)

    def close(self):
        if self.shared is not None:
            self.shared.close()


class OneRankShadow:
    """--check: the whole forest stepped by ONE rank (on this rank's GPU), following the same forest sequence."""

    def __init__(self, run):
        import t8gpu_b200 as tb
        self.tb, self.run = tb, run
        f = run.forest
        lv, cent, vol, _ = f.elements()
        self.sol = self.make(f)
        centres = torch.as_tensor(np.ascontiguousarray(cent.astype(run.npdt))).to(run.device).reshape(-1)
        if run.subgrid:
            from bench_subgrid import cell_centers
            centres = cell_centers(centres, int(lv[0]), run.dtype)
        tb.init_kelvin_helmholtz(3, centres, self.sol.variables(self.sol.next))
        self.steps, self.worst = 0, 0.0

    def make(self, f):
        run, tb = self.run, self.tb
        vol = f.elements()[2].astype(run.npdt)
        if run.subgrid:
            return tb.SubgridEulerSolver(f.connectivity(subgrid=True, dtype=run.npdt), vol, run.dtype, device=run.device,
                                         mode="fused")
        return tb.EulerSolver(f.connectivity(dtype=run.npdt), vol, run.dtype, device=run.device, max_level=run.max_level)

    def volume(self):
        return self.sol.vol if self.run.subgrid else self.sol.volume()

    def iterate(self, dt):
        self.sol.iterate(dt)
        self.steps += 1

    def compare(self):
        run = self.run
        o0, o1 = run.mesh.off[run.rank], run.mesh.off[run.rank + 1]
        S = run.mesh.S
        a = run.state().cpu().numpy().astype(np.float64)
        b = self.sol.state()[:, o0 * S:o1 * S].cpu().numpy().astype(np.float64)
        full = self.sol.state().cpu().numpy().astype(np.float64)
        scale = np.abs(full).max(axis=1)
        scale = np.where(scale < 1e-3 * scale.max(), scale.max(), scale)
        err = float((np.abs(a - b).max(axis=1) / scale).max()) if a.size else 0.0
        same_vol = bool(torch.equal(run.mesh.volume(), self.volume()[o0:o1]))
        self.worst = max(self.worst, err)
        return err, same_vol

    def adapt(self, f_old, f_new):
        tb, run = self.tb, self.run
        amap = f_old.adapt_map(f_new)
        new, old_vol = self.make(f_new), self.volume()
        new_vol = new.vol if run.subgrid else new.volume()
        tb.adapt_remap(torch.as_tensor(amap).to(run.device), self.sol.variables(self.sol.next), new.variables(new.next),
                       old_vol, new_vol, 3 if run.subgrid else 0)
        self.sol = new


def run_amr(args, rank, world, device, dist=None, emit=True):
    """Returns (and on rank 0 prints) the JSON line of the AMR workload."""
    from bench import ClockSampler
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    subgrid = bool(getattr(args, "subgrid", False))
    level = args.level if args.level is not None else (3 if subgrid else 5)
    max_level = level + 2
    every, cycles = args.every, args.cycles
    tol = 1e-12 if args.dtype == "f64" else 1e-5
    run = AmrRun(level, max_level, dtype, rank, world, device, dist, subgrid=subgrid)
    shadow = OneRankShadow(run) if args.check else None
    dt = 0.1 * 2.0 ** -(max_level + (2 if subgrid else 0))
    T = {k: 0.0 for k in ("criteria_device+gather", "forest_host(t8code stand-in)", "adapt_remap_device", "tile_plan",
                          "partition_remap_device")}
    hist, t_step, updates, checks = [], 0.0, 0, []
    for _ in range(3):
        run.iterate(dt)
        if shadow:
            shadow.iterate(dt)
    sampler = ClockSampler(device.index or 0)
    if rank == 0:
        sampler.start()
    wall0 = time.time()
    for cyc in range(cycles):
        n_total = run.forest.num_elements * run.mesh.S
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(every):
            run.iterate(dt)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        t_step += float(ms.item()) * 1e-3
        updates += n_total * every
        if shadow:
            for _ in range(every):
                shadow.iterate(dt)
            err, same_vol = shadow.compare()
            checks.append(dict(cycle=cyc, steps=shadow.steps, rel_linf=err, volumes_bitwise=same_vol,
                               ok=bool(err <= shadow.steps * tol and same_vol)))
        f_old = run.forest
        info, _ = run.adapt(T)
        if shadow:
            shadow.adapt(f_old, run.forest)
            err, same_vol = shadow.compare()       # straight after adapt + partition: both remaps are bit-exact
            checks.append(dict(cycle=cyc, after="adapt+partition", rel_linf=err, volumes_bitwise=same_vol,
                               ok=bool(err <= shadow.steps * tol and same_vol)))
        info["cycle"] = cyc
        hist.append(info)
        assert bool(torch.isfinite(run.state()).all())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.time() - wall0
    clocks = sampler.stop(wall0, time.time()) if rank == 0 else None
    ok = all(c["ok"] for c in checks) if checks else None
    if world > 1:
        flag = torch.tensor([1.0 if ok in (True, None) else 0.0], dtype=torch.float64, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item() == 1.0) if checks else None
    run.mesh.close()
    if run.mail is not None:
        run.mail.close()
    line = {"metric": "cell-updates/s per RK3 step", "value": updates / t_step, "unit": "cell-updates/s", "n_gpus": world,
            "steps": every * cycles, "warmup": 3, "ms_per_step": t_step / (every * cycles) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "kelvin_helmholtz 3D %sadaptive periodic hex forest (levels 1..%d, start uniform level %d), "
                                   "adapt + repartition every %d steps, %d cycles, %d GPU(s)" %
                                   ("Subgrid<4,4,4> on an " if subgrid else "", max_level, level, every, cycles, world),
                       "host_forest": "oracle.Forest: mini-forest stand-in for t8code (absent in this image), replicated "
                                      "on every rank; mesh management only -- stepping, criteria, remaps and plan build "
                                      "are the product",
                       "cycle_seconds": {k: round(v, 4) for k, v in T.items()}, "stepping_seconds": round(t_step, 4),
                       "wall_seconds": round(wall, 3), "end_to_end_cell_updates_per_s": updates / wall,
                       "history": hist, "host_cores": os.cpu_count()},
            "clocks": clocks, "gpu_launches": None,
            "parity": ({"vs": "one-rank run of the same forest sequence on each rank's GPU", "checks": checks,
                        "tolerance_per_step": tol, "ok": ok} if checks else None)}
    if rank == 0 and emit:
        print(json.dumps(line))
    return line


def amr_secondary(dtype_name, rank, world, device, dist=None, level=None, every=10, cycles=2, subgrid=False):
    """Short config-3 (subgrid: config-4) run for the `secondary` block of the default bench lines (with the one-rank
    comparison)."""
    import argparse
    a = argparse.Namespace(level=level, every=every, cycles=cycles, check=True, dtype=dtype_name, subgrid=subgrid)
    line = run_amr(a, rank, world, device, dist, emit=False)
    c, par = line["config"], line["parity"]
    return {"workload": c["workload"], "stepping_cell_updates_per_s": line["value"], "ms_per_step": line["ms_per_step"],
            "end_to_end_cell_updates_per_s": c["end_to_end_cell_updates_per_s"], "cycle_seconds": c["cycle_seconds"],
            "stepping_seconds": c["stepping_seconds"], "history": c["history"], "host_forest": c["host_forest"],
            "parity": {"vs": par["vs"], "ok": par["ok"], "tolerance_per_step": par["tolerance_per_step"],
                       "worst_rel_linf": max(x["rel_linf"] for x in par["checks"]),
                       "volumes_bitwise": all(x["volumes_bitwise"] for x in par["checks"])}}
