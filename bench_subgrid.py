"""`bench.py --workload subgrid`: the Subgrid<4,4,4> path (BASELINE config 4) on 1..8 GPUs, weak scaling.

Each rank owns one tree of a periodic brick of unit trees refined to `--level` (default 6: 262 144 elements x 64 cells =
16 777 216 cells per GPU).  Storage follows SubgridMemoryManager (nb_variables * nb_steps cell arrays of capacity * 64,
per-element volumes apart); the cell arrays live in one cudaIpc-shared allocation per rank, ghost cells are read from
the neighbour GPUs' live arrays through [var][rank] pointer tables, every rank evaluates all faces of its own cells.
The reference's subgrid solver runs with a fixed dt (its compute_timestep body is commented out,
examples/subgrid/solver.inl:309-321), so a step is three stage kernels and three stage barriers, no reduction."""
import json
import os
import time

import torch

ALG_CELL_BYTES = {"f32": 165.8, "f64": 328.3}   # algorithmic bytes per cell per RK3 step, SURVEY.md 8(d) config C4


def subgrid_connectivity(level, dtype, rank, world, device, brick):
    """Uniform periodic brick in the SubgridMeshConnectivityAccessor layout (subgrid_mesh_manager.h:29-216): the element
    level arrays of the device builder plus, for same-level faces, level difference 0 and the anchor cell of the face
    inside the right element (subgrid_mesh_manager.inl:590-631: 0 behind a +axis face, 3 behind a -axis face)."""
    import t8gpu_b200 as tb
    conn = tb.cartesian_uniform_connectivity(3, level, dtype, world, rank, device=device, brick=brick)

    def anchors(normals):
        nrm = normals.view(-1, 3)
        ax = nrm.abs().argmax(1, keepdim=True)
        off = torch.zeros(nrm.shape, dtype=torch.int32, device=nrm.device)
        off.scatter_(1, ax, torch.where(nrm.gather(1, ax) > 0, 0, 3).to(torch.int32))
        return off.reshape(-1)

    conn["offsets"] = anchors(conn["face_normals"])
    conn["level_diff"] = torch.zeros(int(conn["n_faces"]), dtype=torch.int32, device=device)
    conn["x_offsets"] = anchors(conn["x_face_normals"])
    conn["x_level_diff"] = torch.zeros(int(conn["n_xfaces"]), dtype=torch.int32, device=device)
    return conn


def cell_centers(centroids, level, dtype):
    """Centres of the 4x4x4 cells of every element (examples/subgrid/solver.inl:13-35), cell index i + 4j + 16k."""
    h = 0.5 ** level
    c = centroids.view(-1, 1, 3).to(torch.float64)
    i = torch.arange(64, device=c.device)
    ijk = torch.stack([i % 4, (i // 4) % 4, i // 16], 1).to(torch.float64)
    return (c - 0.5 * h + 0.125 * h + ijk.view(1, 64, 3) * (0.25 * h)).to(dtype).reshape(-1)


class SubgridBrick:
    def __init__(self, level, dtype, rank, world, device, brick=None):
        import t8gpu_b200 as tb
        import torch.distributed as dist
        from t8gpu_b200.multi import BRICK, exchange_wires, row_pointers
        from t8gpu_b200.solver import NB_STEPS, NVAR
        self.rank, self.world, self.device, self.dtype = rank, world, device, dtype
        esz = 8 if dtype == torch.float64 else 4
        conn = subgrid_connectivity(level, dtype, rank, world, device, brick or BRICK[world])
        self.ne = int(conn["n_local"])
        self.n = self.ne * 64
        self.n_faces = int(conn["n_faces"]) + int(conn["n_xfaces"])
        self.n_ghost = int(conn["n_ghost"])
        self.vol = conn["volumes"]
        self.sync = os.environ.get("T8B200_SYNC", "push") if world > 1 else "single"
        tail = self.sync in ("push", "pull")
        t0 = time.time()
        self.plan = tb.SubgridPlan.from_device(conn, self.vol, dtype, ghost_tail=tail)   # no D2H of the connectivity
        if self.plan is None:
            self.plan = tb.SubgridPlan(tb.conn_to_host(conn), self.vol.cpu().numpy(), dtype, ghost_tail=tail)
        torch.cuda.synchronize()
        self.t_plan = time.time() - t0
        self.cap = (self.n + self.plan.n_tail + 31) // 32 * 32     # own cells, then the ghost-cell tail
        self.shared = tb.SharedBuffer(NVAR * NB_STEPS * self.cap * esz, device)
        self.buffer = self.shared.tensor((NVAR * NB_STEPS, self.cap), dtype)
        self.tables, self.mail = {s: None for s in range(NB_STEPS)}, None
        if world > 1:
            wires = exchange_wires(dist, self.shared.handle, self.cap, world, device)
            bases = [self.shared.ptr if r == rank else self.shared.open_peer(wires[r][0]) for r in range(world)]
            rows = row_pointers(bases, [c for _, c in wires], NVAR, NB_STEPS, esz)
            self.tables = {s: tb.PointerTables(rows[s], device) for s in range(NB_STEPS)}
            self.mail = tb.PeerMailboxes(rank, world, device)
            self.mail.exchange([h for h, _ in exchange_wires(dist, self.mail.handle, 0, world, device)])
            if self.sync == "push":        # what the peers hold copies of, sorted by destination (one all-gather)
                from t8gpu_b200.multi import send_lists
                self.send = send_lists(dist, self.plan, self.n, rank, world, device)
        tb.init_kelvin_helmholtz(3, cell_centers(conn["centroids"], level, dtype), self.variables(0))
        self.next, self.prev = 0, 3
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            self.publish()

    def publish(self):
        """State written outside the stage kernels: full barrier before any rank reads ghosts (starts / continues the
        stage-epoch sequence of the self-ordering kernels)."""
        if self.mail is not None:
            if self.sync == "push":
                import t8gpu_b200 as tb
                tb.ghost_push(*self.send, self.variables(self.next), self.tables[self.next])
            self.mail.barrier()

    def variables(self, step):
        return [self.buffer[step * 5 + k, :self.n] for k in range(5)]

    def state(self):
        return self.buffer[self.next * 5:(self.next + 1) * 5, :self.n]

    def iterate(self, dt):
        self.next, self.prev = self.prev, self.next
        prev, s1, s2, nxt = (self.variables(s) for s in (self.prev, 1, 2, self.next))
        T, launches = self.tables, 0
        own_order = self.mail is not None and self.sync == "kernel"   # the stage kernels order themselves
        pull = self.mail is not None and self.sync == "pull"          # ghost-cell tail, single-rank stage kernels
        push = self.mail is not None and self.sync == "push"          # ... filled by the owner right behind the stage
        for stage, sin, sout, vin, vout in ((1, self.prev, 1, prev, s1), (2, 1, 2, s1, s2), (3, 2, self.next, s2, nxt)):
            if pull:
                self.plan.pull(vin, T[sin])
                launches += 1
            self.plan.stage(stage, vin, prev if stage > 1 else None, vout, self.vol, dt,
                            in_all=None if (pull or push) else T[sin], sync=self.mail if own_order else None)
            launches += 1
            if push:
                import t8gpu_b200 as tb
                tb.ghost_push(*self.send, vout, T[sout])
                launches += 1
            if self.mail is not None and not own_order:   # barrier kernel per stage
                self.mail.barrier()
                launches += 1
        return launches

    def close(self):
        torch.cuda.synchronize()
        self.plan = None
        if self.mail is not None:
            import torch.distributed as dist
            dist.barrier()
            self.mail.close()
        self.shared.close()


def parity_subgrid(sol, level, dtype_name, steps=3):
    """The measured configuration against the reference's own subgrid kernels (oracle/_ref, the checker) from the same
    Kelvin-Helmholtz cell state; relative L-infinity after each step."""
    import numpy as np
    tol = 1e-12 if dtype_name == "f64" else 1e-5
    npdt = np.float64 if dtype_name == "f64" else np.float32
    try:
        from oracle import ref_cuda
        if not ref_cuda.available():
            raise RuntimeError
    except Exception:
        return {"vs": "unavailable (oracle/_ref not built)", "ok": None}
    t0 = time.time()
    dt = 0.1 * 2.0 ** -(level + 2)
    u0 = sol.state().cpu().numpy().astype(npdt)
    ref = ref_cuda.RefSolver("sg", npdt, 3, level, True)
    ref.set_state(u0)
    errs = []
    for _ in range(steps):
        ref.iterate(dt)
        sol.iterate(dt)
        a, b = sol.state().cpu().numpy().astype(np.float64), ref.get_state().astype(np.float64)
        scale = np.abs(b).max(axis=1)
        scale = np.where(scale < 1e-3 * scale.max(), scale.max(), scale)
        errs.append(float((np.abs(a - b).max(axis=1) / scale).max()))
    ref.close()
    return {"vs": "reference CUDA kernels (oracle/_ref, examples/subgrid compiled unmodified)", "level": level,
            "cells": int(u0.shape[1]), "steps": steps, "rel_linf_after_step": errs, "tolerance_per_step": tol,
            "ok": all(e <= (k + 1) * tol for k, e in enumerate(errs)), "seconds": round(time.time() - t0, 1)}


def measure_subgrid(level, dtype_name, steps, warmup, device, parity=True):
    """Short single-GPU Subgrid<4,4,4> run for the `secondary` block of the default bench line."""
    from bench import ClockSampler, measured_peak
    dtype = torch.float64 if dtype_name == "f64" else torch.float32
    sol = SubgridBrick(level, dtype, 0, 1, device)
    dt = 0.1 * 2.0 ** -(level + 2)
    stream = torch.cuda.current_stream()
    for _ in range(warmup):
        sol.iterate(dt)
    torch.cuda.synchronize()
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record(stream)
    for _ in range(steps):
        sol.iterate(dt)
    ev1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop(t0, time.time())
    ms_per_step = ev0.elapsed_time(ev1) / steps
    assert bool(torch.isfinite(sol.state()).all()), "the run diverged"
    peak, _ = measured_peak()
    achieved = ALG_CELL_BYTES[dtype_name] * sol.n / (ms_per_step * 1e-3) / 1e9
    out = {"workload": "Subgrid<4,4,4>, uniform periodic hex forest level %d (%d cells) %s" % (level, sol.n, dtype_name),
           "ms_per_step": ms_per_step, "value": sol.n / (ms_per_step * 1e-3), "clocks": clocks,
           "roofline_frac": achieved / peak, "roofline_achieved_gbs": achieved,
           "alg_bytes_per_cell_step": ALG_CELL_BYTES[dtype_name], "tile_plan_s": round(sol.t_plan, 2)}
    if dtype_name == "f64":
        # the FP64-pipe roofline beside the HBM one: the structured kernel evaluates the same 3.5 faces per cell as per
        # element of the unstructured mesh (same template, SubgridBox layout), so the executed FP64 instructions per
        # cell-stage are those of the committed capture of that kernel
        try:
            import json
            from bench import ROOT, fp64_peak
            rec = json.load(open(os.path.join(ROOT, "profiles", "traffic_f64.json")))
            per_cell = rec["fp64_thread_inst_per_launch"] / rec.get("elements_per_launch", 16777216)
            pk, pk_src = fp64_peak()
            rate = per_cell * sol.n / (ms_per_step / 3.0 * 1e-3)
            out["fp64_pipe"] = {"frac": rate / pk, "per_cell_stage": per_cell, "peak": pk, "peak_source": pk_src,
                                "floor_ms_per_step": 3e3 * per_cell * sol.n / pk,
                                "note": "HBM roofline counts state bytes only (328.3 B per cell-step): its FP64-pipe "
                                        "ceiling sits at %.0f %% of that roofline" % (100 * ALG_CELL_BYTES["f64"] * sol.n / peak / 1e9 / (3e3 * per_cell * sol.n / pk * 1e-3))}
        except (OSError, KeyError, ValueError):
            pass
    if parity:
        out["parity"] = parity_subgrid(sol, level, dtype_name)
    sol.close()
    return out


def bitwise_parity(dtype, rank, world, device, level=2, steps=3):
    """The N-rank Subgrid<4,4,4> run must be bitwise the one-rank run of the same brick (see bench_multigpu)."""
    import torch.distributed as dist
    from t8gpu_b200.multi import BRICK
    dt = 0.1 * 2.0 ** -(level + 2)
    sol = SubgridBrick(level, dtype, rank, world, device)
    sub = torch.cuda.Stream()            # the one-rank brick needs no collectives: built with world = 1
    one = SubgridBrick(level, dtype, 0, 1, device, brick=BRICK[world])
    n1 = one.n
    off = [(n1 // 64) * r // world * 64 for r in range(world + 1)]
    same = bool(torch.equal(one.state()[:, off[rank]:off[rank + 1]], sol.state()))
    for _ in range(steps):
        one.iterate(dt)
        sol.iterate(dt)
    torch.cuda.synchronize()
    same = same and bool(torch.equal(one.state()[:, off[rank]:off[rank + 1]], sol.state()))
    flag = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    one.close()
    sol.close()
    del sub
    return {"vs": "one-rank run of the same brick (level %d trees, %d cells)" % (level, n1), "steps": steps,
            "bitwise_equal": bool(flag.item() == 1.0)}


def run_subgrid(args, rank, world, device):
    import torch.distributed as dist
    from bench import ClockSampler, measured_peak
    from t8gpu_b200.multi import BRICK
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    level = args.level

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    parity = None
    if world > 1:
        parity = bitwise_parity(dtype, rank, world, device)
        assert parity["bitwise_equal"], "the %d-rank subgrid run differs from the one-rank run of the same brick" % world
    t0 = time.time()
    sol = SubgridBrick(level, dtype, rank, world, device)
    t_setup = time.time() - t0
    dt = 0.1 * 2.0 ** -(level + 2)           # examples/subgrid/main_3d.cu:27-30
    n, stream = sol.n, torch.cuda.current_stream()
    for _ in range(args.warmup):
        sol.iterate(dt)
    barrier()
    sampler = ClockSampler(device.index)
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
    tw1 = time.time()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    assert bool(torch.isfinite(sol.state()).all()), "the run diverged"

    # e2e: pinned-host state in, K x iterate (each step ends with a stream synchronisation, the reference's
    # cudaDeviceSynchronize at the end of iterate(), solver.inl:264), state out
    u_host = torch.empty((5, n), dtype=dtype).pin_memory()
    u_host.copy_(sol.state())
    out_host = torch.empty((5, n), dtype=dtype).pin_memory()
    probe = torch.empty(1, dtype=dtype).pin_memory()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sol.state().copy_(u_host, non_blocking=True)
    sol.publish()                              # every rank's upload is complete before neighbours read it
    for _ in range(args.steps):
        sol.iterate(dt)
        probe.copy_(sol.state()[0, :1], non_blocking=True)
        stream.synchronize()
    out_host.copy_(sol.state(), non_blocking=True)
    e1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    info, nfaces, nghost, ne, t_plan = sol.plan.info, sol.n_faces, sol.n_ghost, sol.ne, sol.t_plan
    sol.close()
    if rank == 0:
        esz = 8 if args.dtype == "f64" else 4
        total, ms_per_step = n * world, ms / args.steps
        peak, src = measured_peak()
        alg = ALG_CELL_BYTES[args.dtype]
        achieved = alg * n / (ms_per_step * 1e-3) / 1e9     # per GPU
        state_bytes = 5 * n * esz
        line = {"metric": "cell-updates/s per RK3 step", "value": total * args.steps / (ms * 1e-3),
                "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": "kelvin_helmholtz 3D Subgrid<4,4,4> on a uniform periodic hex forest, brick %s of "
                                       "level-%d trees (%d elements = %d cells per GPU, %d cells total) %s, fixed dt, no "
                                       "adaptation" % (BRICK[world], level, ne, n, total, args.dtype),
                           "cells_per_gpu": n, "element_faces_per_gpu": nfaces, "ghost_elements_per_gpu": nghost,
                           "partition": "one tree per GPU; ghost cells read from peer GPUs over NVLink, owner-computes "
                                        "boundary faces; per stage the single-rank stage kernel, an owner-side push of the "
                                        "ghost-cell copies, one mailbox barrier kernel" if world > 1 else
                                        "one rank", "l2": "inputs larger than L2 (%.0f MB of state per stage)" %
                                                          (2 * state_bytes / 1e6),
                           "host_setup_s": round(t_setup, 2), "tile_plan_host_s": round(t_plan, 2),
                           "host_cores": os.cpu_count(), "plan": info},
                "clocks": clocks, "parity": parity,
                "e2e": {"value": total * args.steps / (e2e_ms * 1e-3), "unit": "cell-updates/s",
                        "h2d_bytes_per_step": state_bytes / args.steps * world,
                        "d2h_bytes_per_step": (state_bytes / args.steps + esz) * world,
                        "ms_per_step": e2e_ms / args.steps,
                        "protocol": "pinned-host state in, K x (iterate + D2H probe + stream sync), state out"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": src,
                             "kernel": "fused_stage_kernel (cell-level tile plan)", "note": "per GPU",
                             "alg_bytes_per_launch": alg * n / 3.0, "avg_launch_ms": ms_per_step / 3.0}}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.dtype)
        print(json.dumps(line))


def cpu_baseline(dtype_name, budget_s=10.0, level=3):
    """The CPU oracle (kind "port": the reference has no CPU implementation of this path) on a bounded sample of the
    same workload: Subgrid<4,4,4> on the uniform periodic hex forest of level 3 (32 768 cells), single thread."""
    import numpy as np
    import oracle
    npdt = np.float64 if dtype_name == "f64" else np.float32
    f = oracle.Forest(3, level)
    conn = f.connectivity(subgrid=True, dtype=npdt)
    lv, cent, vol, _ = f.elements()
    u = oracle.subgrid_init_kh(3, cent.astype(npdt), lv, npdt)
    vol = vol.astype(npdt)
    dt = 0.1 * 2.0 ** -(level + 2)
    steps, t0 = 0, time.time()
    while steps < 2 or time.time() - t0 < budget_s:
        u, _, _ = oracle.subgrid_iterate(conn, vol, u, dt)
        steps += 1
    el = time.time() - t0
    return {"value": u.shape[1] * steps / el, "unit": "cell-updates/s", "cores": 1, "kind": "port",
            "sample": "Subgrid<4,4,4>, uniform periodic hex level %d (%d cells), %d RK3 steps, %.1f s, %s" %
                      (level, u.shape[1], steps, el, dtype_name)}
