"""Multi-GPU arm of bench.py: one process per GPU (torchrun), weak scaling.

Each rank owns one tree of a periodic brick of unit trees refined to `--level` (16.8 M hexes per GPU at level 8):
brick (2,1,1) / (2,2,1) / (2,2,2) for 2 / 4 / 8 GPUs.  Ghost states are read directly from the neighbour GPUs'
live state arrays through peer-mapped pointers (NVLink); every rank evaluates all faces of its own elements
(owner computes, no remote atomics).  One tiny NCCL all-reduce per RK stage orders the stages across ranks on the
device (no host synchronisation inside a step); the stage-3 one carries the max wave speed (CFL reduction)."""
import json
import os
import time

import torch
import torch.distributed as dist

from t8gpu_b200.multi import BRICK, exchange_wires, global_max_wave_speed, row_pointers, stage_barrier, timestep


class MultiGpuEuler:
    def __init__(self, level, dtype, rank, world, device, brick=None):
        import t8gpu_b200 as tb
        from t8gpu_b200.solver import NB_STEPS, NVAR
        self.tb, self.rank, self.world, self.device, self.dtype = tb, rank, world, device, dtype
        brick = brick or BRICK[world]
        esz = 8 if dtype == torch.float64 else 4
        conn = tb.cartesian_uniform_connectivity(3, level, dtype, world, rank, device=device, brick=brick)
        self.n = int(conn["n_local"])
        self.n_faces = int(conn["n_faces"]) + int(conn["n_xfaces"])
        self.n_ghost = int(conn["n_ghost"])
        self.cap = (self.n + 31) // 32 * 32
        nrows = NVAR * NB_STEPS + 1
        self.shared = tb.SharedBuffer(nrows * self.cap * esz, device)
        self.buffer = self.shared.tensor((nrows, self.cap), dtype)
        self.buffer[NVAR * NB_STEPS, :self.n] = conn["volumes"]
        # exchange (handle, capacity) with every rank, map the peers, fill the [var][rank] tables
        wires = exchange_wires(dist, self.shared.handle, self.cap, world, device)
        caps = [c for _, c in wires]
        bases = [self.shared.ptr if r == rank else self.shared.open_peer(wires[r][0]) for r in range(world)]
        rows = row_pointers(bases, caps, NVAR, NB_STEPS, esz)
        self.tables = {s: tb.PointerTables(rows[s], device) for s in range(NB_STEPS)}
        # stage barrier + CFL max over peer memory (T8B200_SYNC=nccl: 1-element NCCL all-reduces instead)
        self.mail = None
        if os.environ.get("T8B200_SYNC", "peer") == "peer":
            self.mail = tb.PeerMailboxes(rank, world, device)
            mw = exchange_wires(dist, self.mail.handle, 0, world, device)
            self.mail.exchange([h for h, _ in mw])
            self.speed_loc = torch.zeros(1, dtype=dtype, device=device)
        self.plan = tb.Plan(tb.conn_to_host(conn), dtype)
        tb.init_kelvin_helmholtz(3, conn["centroids"], self.variables(0))
        self.next, self.prev = 0, 3
        self.speed_max = torch.zeros(1, dtype=dtype, device=device)
        self.token = torch.zeros(1, dtype=dtype, device=device)
        torch.cuda.synchronize()
        dist.barrier()

    def variables(self, step):
        from t8gpu_b200.solver import NVAR
        return [self.buffer[step * NVAR + k, :self.n] for k in range(NVAR)]

    def volume(self):
        from t8gpu_b200.solver import NB_STEPS, NVAR
        return self.buffer[NVAR * NB_STEPS, :self.n]

    def iterate(self, dt):
        self.next, self.prev = self.prev, self.next
        prev, s1, s2, nxt = (self.variables(s) for s in (self.prev, 1, 2, self.next))
        vol, T = self.volume(), self.tables
        if self.mail is not None:   # barriers and the CFL reduction through the peers' mailboxes (NVLink stores)
            self.plan.stage(1, prev, None, s1, vol, dt, in_all=T[self.prev])
            self.mail.barrier()
            self.plan.stage(2, s1, prev, s2, vol, dt, in_all=T[1])
            self.mail.barrier()
            self.plan.stage(3, s2, prev, nxt, vol, dt, in_all=T[2], speed_max=self.speed_loc)
            self.mail.barrier(self.speed_loc, self.speed_max)
            return 6                                               # 3 stage kernels + 3 barrier kernels
        self.plan.stage(1, prev, None, s1, vol, dt, in_all=T[self.prev])
        stage_barrier(dist, self.token)                            # device-side, on the compute stream
        self.plan.stage(2, s1, prev, s2, vol, dt, in_all=T[1])
        stage_barrier(dist, self.token)
        self.plan.stage(3, s2, prev, nxt, vol, dt, in_all=T[2], speed_max=self.speed_max)
        global_max_wave_speed(dist, self.speed_max)                # barrier + global CFL reduction
        return 3

    def close(self):
        torch.cuda.synchronize()
        dist.barrier()
        self.plan = None
        if self.mail is not None:
            self.mail.close()
        self.shared.close()


def run_multi(args, rank, world, device):
    from bench import ALG_BYTES, ClockSampler, measured_peak
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    t0 = time.time()
    sol = MultiGpuEuler(args.level, dtype, rank, world, device)
    t_setup = time.time() - t0
    dt = 0.1 * 2.0 ** -args.level
    n = sol.n
    stream = torch.cuda.current_stream()
    for _ in range(args.warmup):
        sol.iterate(dt)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(device.index)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    tw0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        launches += sol.iterate(dt)
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    tw1 = time.time()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    vmax = float(sol.speed_max.item())
    assert vmax > 0 and vmax == vmax

    # e2e: same protocol as the 1-GPU arm (state upload from pinned host, per-step D2H of the reduced max wave speed
    # with the next dt computed on the host, state download), max over ranks
    u_host = torch.empty((5, n), dtype=dtype).pin_memory()
    state = sol.buffer[sol.next * 5:(sol.next + 1) * 5, :n]
    u_host.copy_(state)
    out_host = torch.empty((5, n), dtype=dtype).pin_memory()
    vmax_host = torch.empty(1, dtype=dtype).pin_memory()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sol.buffer[sol.next * 5:(sol.next + 1) * 5, :n].copy_(u_host, non_blocking=True)
    dist.all_reduce(sol.token, op=dist.ReduceOp.MAX)   # every rank's upload is complete before neighbours read it
    cur_dt = dt
    for _ in range(args.steps):
        sol.iterate(cur_dt)
        vmax_host.copy_(sol.speed_max, non_blocking=True)
        stream.synchronize()
        cur_dt = timestep(float(vmax_host[0]), 0.7, args.level, dt_cap=dt)
    out_host.copy_(sol.buffer[sol.next * 5:(sol.next + 1) * 5, :n], non_blocking=True)
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())

    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    info = sol.plan.info
    nfaces, nghost = sol.n_faces, sol.n_ghost
    sol.close()
    if rank == 0:
        esz = 8 if args.dtype == "f64" else 4
        total = n * world
        peak, src = measured_peak()
        alg = ALG_BYTES[("hex", args.dtype)]
        ms_per_step = ms / args.steps
        achieved = alg * n / (ms_per_step * 1e-3) / 1e9   # per GPU
        state_bytes = 5 * n * esz
        line = {"metric": "cell-updates/s per RK3 step", "value": total * args.steps / (ms * 1e-3),
                "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": "kelvin_helmholtz 3D uniform periodic hex mesh, brick %s of level-%d trees "
                                       "(%d elements per GPU, %d total) %s, fixed dt, no adaptation" %
                                       (BRICK[world], args.level, n, total, args.dtype),
                           "elements_per_gpu": n, "faces_per_gpu": nfaces, "ghosts_per_gpu": nghost,
                           "partition": "one tree (one t8code SFC partition) per GPU; ghosts read from peer GPUs "
                                        "over NVLink (cudaIpc-mapped state arrays), owner-computes boundary faces",
                           "sync": ("peer-memory mailbox barrier per RK stage on the compute stream (NVLink stores + "
                                    "acquire spin); stage 3 carries the max wave speed" if os.environ.get(
                                        "T8B200_SYNC", "peer") == "peer" else
                                    "1 NCCL all-reduce (1 element) per RK stage on the compute stream; stage 3 "
                                    "carries the max wave speed"), "l2": "inputs larger than L2",
                           "host_setup_s": round(t_setup, 2), "host_cores": os.cpu_count(), "plan": info},
                "clocks": clocks,
                "e2e": {"value": total * args.steps / (e2e_ms * 1e-3), "unit": "cell-updates/s",
                        "h2d_bytes_per_step": (state_bytes / args.steps + esz) * world,
                        "d2h_bytes_per_step": (state_bytes / args.steps + esz) * world,
                        "ms_per_step": e2e_ms / args.steps,
                        "protocol": "pinned-host state in, K x (iterate + D2H max wave speed + host dt), state out"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": src,
                             "kernel": "fused_stage_kernel", "note": "per GPU"},
                "max_wave_speed": vmax}
        print(json.dumps(line))
    dist.destroy_process_group()
