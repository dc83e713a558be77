"""Multi-GPU arm of bench.py: one process per GPU (torchrun), weak scaling.

Each rank owns one tree of a periodic brick of unit trees refined to `--level` (16.8 M hexes per GPU at level 8):
brick (2,1,1) / (2,2,1) / (2,2,2) for 2 / 4 / 8 GPUs.  Ghost states are read directly from the neighbour GPUs'
live state arrays through peer-mapped pointers (NVLink); every rank evaluates all faces of its own elements
(owner computes, no remote atomics).

Stage ordering across ranks (T8B200_SYNC):
  push (default)    ghost tail as in `pull`, filled from the OTHER side: after each stage every rank pushes the elements
                    its peers hold copies of into their tails (t8b200_ghost_push: scattered local gather, consecutive
                    remote stores), then the mailbox barrier publishes them; the next stage kernel starts right behind
                    the barrier.
  pushb             push and barrier in ONE launch (t8b200_ghost_push_barrier: the last CTA to finish its pushes signals
                    and waits): two launches per stage instead of three.
  fpush             the push folded into the stage kernel (t8b200_fused_stage_push): the thread that writes an element
                    its peers hold a copy of also stores the new values into those copies from the kernel's epilogue.
                    Bitwise equal, but measured SLOWER than the separate push kernel on 2 GPUs (2.57 vs 2.46 ms per
                    step): the remote stores of a warp are isolated 8-byte packets, the push kernel sorts them by
                    destination into whole lines.  Kept as an option.
  pull              ghost tail: every rank keeps local copies of its ghosts behind its own elements; per stage one
                    mailbox barrier kernel, one bandwidth-bound pull kernel that copies the ghosts from the peers' rows
                    over NVLink (no pack / unpack on the owner's side), then the SINGLE-rank stage kernel, which reads
                    no peer memory (direct peer loads inside the stage kernel cost +8 % per step: the NVLink latency
                    is exposed in every chunk at the partition boundary).  The stage-3 barrier carries the CFL max.
  overlap           ghost tail as in `pull`, with every stage launched in two passes (t8b200_fused_stage_part): the
                    chunks that read no ghost copy on the compute stream right behind the previous stage, while a
                    second (high-priority) stream does barrier -> pull -> the partition-boundary chunks.  Measured
                    SLOWER than `pull` on 2 GPUs (2.61 vs 2.50 ms per step): the boundary pass does not fill the GPU and
                    the interior pass of the next stage has to wait for it.  Kept as an option.
  kernel            direct peer loads; the stage kernels order themselves: the chunks that read ghost elements run first, wait for the
                    peers' previous stage and the last of them signals this stage to every peer through the
                    peer-mapped mailboxes (csrc/peer_sync.cuh); all other chunks never wait.  The CFL reduction is one
                    32-thread mailbox kernel per step (on a side stream while dt is fixed, on the compute stream when
                    the next dt depends on it).
  peer              one mailbox barrier kernel per stage (round-1 scheme)
  nccl              one 1-element NCCL all-reduce per stage
No host synchronisation inside a step in any mode."""
import json
import os
import time

import torch
import torch.distributed as dist

from t8gpu_b200.multi import BRICK, exchange_wires, global_max_wave_speed, row_pointers, stage_barrier

CFL = 0.7   # examples/compressible_euler/solver.h:37


def pin_to_gpu_numa_node(device):
    """Host threads onto the CPUs next to this rank's GPU, and the memory policy of the process to the GPU's NUMA node
    (set_mempolicy(MPOL_PREFERRED)): the pinned staging buffers allocated afterwards sit next to the GPU even when the
    container owns no CPU of that socket.  Returns (local cpus bound, numa node or -1, policy set?)."""
    ncpu, node, policy = 0, -1, False
    try:
        props = torch.cuda.get_device_properties(device)
        base = "/sys/bus/pci/devices/%04x:%02x:%02x.0/" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        cpus = set()
        for part in open(base + "local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            ncpu = len(cpus)
        node = int(open(base + "numa_node").read().strip())
    except Exception:
        pass
    if node >= 0 and os.environ.get("T8B200_MEMPOLICY", "1") != "0":
        try:
            import ctypes
            import platform
            if platform.machine() == "x86_64":
                mask = (ctypes.c_ulong * 16)()
                mask[node // 64] = 1 << (node % 64)
                libc = ctypes.CDLL(None, use_errno=True)
                policy = libc.syscall(238, 1, mask, 1024) == 0        # SYS_set_mempolicy, MPOL_PREFERRED
        except Exception:
            policy = False
    return ncpu, node, policy


class MultiGpuEuler:
    def __init__(self, level, dtype, rank, world, device, brick=None, sync=None):
        import t8gpu_b200 as tb
        from t8gpu_b200.solver import NB_STEPS, NVAR
        self.tb, self.rank, self.world, self.device, self.dtype = tb, rank, world, device, dtype
        self.sync = sync or os.environ.get("T8B200_SYNC", "push")
        brick = brick or BRICK[world]
        esz = 8 if dtype == torch.float64 else 4
        conn = tb.cartesian_uniform_connectivity(3, level, dtype, world, rank, device=device, brick=brick)
        self.n = int(conn["n_local"])
        self.n_faces = int(conn["n_faces"]) + int(conn["n_xfaces"])
        self.n_ghost = int(conn["n_ghost"])
        t0 = time.time()
        self.plan = tb.Plan.from_device(conn, dtype, ghost_tail=self.sync in ("fpush", "push", "pushb", "pull", "overlap"))     # no D2H
        if self.plan is None:
            self.plan = tb.Plan(tb.conn_to_host(conn), dtype, ghost_tail=self.sync in ("fpush", "push", "pushb", "pull", "overlap"))
        torch.cuda.synchronize()
        self.t_plan = time.time() - t0
        self.cap = (self.n + self.plan.n_tail + 31) // 32 * 32     # own elements, then the ghost tail
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
        self.mail = None
        if self.sync in ("fpush", "push", "pushb", "overlap", "pull", "kernel", "peer", "none"):
            self.mail = tb.PeerMailboxes(rank, world, device)
            mw = exchange_wires(dist, self.mail.handle, 0, world, device)
            self.mail.exchange([h for h, _ in mw])
        self.speed_loc = torch.zeros(1, dtype=dtype, device=device)
        if self.sync in ("push", "fpush", "pushb"):
            from t8gpu_b200.multi import send_csr, send_lists
            self.send = send_lists(dist, self.plan, self.n, rank, world, device)
            self.csr = send_csr(self.send, self.n, device)
        tb.init_kelvin_helmholtz(3, conn["centroids"], self.variables(0))
        self.next, self.prev = 0, 3
        self.speed_max = torch.zeros(1, dtype=dtype, device=device)
        self.token = torch.zeros(1, dtype=dtype, device=device)
        self.dt_dev = torch.zeros(1, dtype=dtype, device=device)
        self.length = 0.5 ** level                # the length scale of compute_timestep (solver.cu:225-228)
        # high priority: the boundary pass must get SM slots ahead of the interior pass' remaining chunks, or it would
        # run at the tail of the stage and the next interior pass (which needs it) could not start
        self.side = torch.cuda.Stream(device=device, priority=-1)
        self.side_done = None
        self.ev_b = None
        torch.cuda.synchronize()
        dist.barrier()
        self.publish()

    def publish(self):
        """The state was written outside the stage kernels (initial data, upload): a full barrier before any rank reads
        ghosts, which also starts the stage-epoch sequence the self-ordering kernels continue."""
        if self.mail is not None and self.sync != "none":
            if self.sync in ("push", "fpush", "pushb"):     # the state was written by other means: push its ghost copies, then publish
                self.tb.ghost_push(*self.send, self.variables(self.next), self.tables[self.next])
                self.mail.barrier(self.speed_loc, self.speed_max)
            elif self.sync in ("pull", "overlap"):    # same class as the barrier that ends a step (precedes the first pull)
                self.mail.barrier(self.speed_loc, self.speed_max)
                if self.sync == "overlap":          # the second stream continues from here
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream())
                    self.side.wait_event(ev)
                    self.ev_b = None
            else:
                self.mail.barrier()

    def variables(self, step):
        from t8gpu_b200.solver import NVAR
        return [self.buffer[step * NVAR + k, :self.n] for k in range(NVAR)]

    def volume(self):
        from t8gpu_b200.solver import NB_STEPS, NVAR
        return self.buffer[NVAR * NB_STEPS, :self.n]

    def state(self):
        return self.buffer[self.next * 5:(self.next + 1) * 5, :self.n]

    def iterate(self, dt, adaptive=False):
        """One RK3 step.  adaptive: dt is read from self.dt_dev (device) and the next dt = min(dt, cfl h / vmax) is
        written there after the global CFL reduction (compute_timestep, solver.cu:213-229) -- no host round trip."""
        self.next, self.prev = self.prev, self.next
        prev, s1, s2, nxt = (self.variables(s) for s in (self.prev, 1, 2, self.next))
        vol, T = self.volume(), self.tables
        dtd = self.dt_dev if adaptive else None
        if self.sync == "fpush":
            m, P = self.mail, self.plan
            ok = P.stage_push(1, prev, None, s1, T[1], vol, dt, self.csr, dt_dev=dtd)
            if not ok:                                    # not a structured-only plan: separate push kernel
                self.sync = "push"
                self.next, self.prev = self.prev, self.next
                return self.iterate(dt, adaptive)
            m.barrier()
            P.stage_push(2, s1, prev, s2, T[2], vol, dt, self.csr, dt_dev=dtd)
            m.barrier()
            P.stage_push(3, s2, prev, nxt, T[self.next], vol, dt, self.csr, speed_max=self.speed_loc, dt_dev=dtd)
            m.barrier(self.speed_loc, self.speed_max)     # stage barrier + CFL max over the ranks
            if adaptive:
                self.tb.timestep(self.speed_max, CFL, self.length, dt, self.dt_dev)
            return 6 + int(adaptive)                      # 3 x (stage with push, barrier)
        if self.sync == "pushb":
            m, P = self.mail, self.plan
            P.stage(1, prev, None, s1, vol, dt, dt_dev=dtd)
            m.push_barrier(*self.send, s1, T[1])          # push + the barrier that publishes it, one launch
            P.stage(2, s1, prev, s2, vol, dt, dt_dev=dtd)
            m.push_barrier(*self.send, s2, T[2])
            P.stage(3, s2, prev, nxt, vol, dt, speed_max=self.speed_loc, dt_dev=dtd)
            m.push_barrier(*self.send, nxt, T[self.next], self.speed_loc, self.speed_max)   # + CFL max over the ranks
            if adaptive:
                self.tb.timestep(self.speed_max, CFL, self.length, dt, self.dt_dev)
            return 6 + int(adaptive)                      # 3 x (stage, push + barrier)
        if self.sync == "push":
            m, P, push = self.mail, self.plan, self.tb.ghost_push
            P.stage(1, prev, None, s1, vol, dt, dt_dev=dtd)
            push(*self.send, s1, T[1])                    # the peers' copies of this rank's boundary elements
            m.barrier()
            P.stage(2, s1, prev, s2, vol, dt, dt_dev=dtd)
            push(*self.send, s2, T[2])
            m.barrier()
            P.stage(3, s2, prev, nxt, vol, dt, speed_max=self.speed_loc, dt_dev=dtd)
            push(*self.send, nxt, T[self.next])
            m.barrier(self.speed_loc, self.speed_max)     # stage barrier + CFL max over the ranks
            if adaptive:
                self.tb.timestep(self.speed_max, CFL, self.length, dt, self.dt_dev)
            return 9 + int(adaptive)                      # 3 x (stage, push, barrier)
        if self.sync == "overlap":
            m, P, main, side = self.mail, self.plan, torch.cuda.current_stream(), self.side
            for stage, sin, vin, vout in ((1, self.prev, prev, s1), (2, 1, s1, s2), (3, 2, s2, nxt)):
                sm = self.speed_loc if stage == 3 else None
                pv = prev if stage > 1 else None
                # compute stream: the chunks without ghosts, as soon as this rank's previous stage is complete
                if self.ev_b is not None:
                    main.wait_event(self.ev_b)
                if not P.stage_part(stage, 1, vin, pv, vout, vol, dt, speed_max=sm, dt_dev=dtd):
                    self.sync = "pull"                    # not a structured-only plan: serial scheme from here on
                    self.next, self.prev = self.prev, self.next
                    main.wait_stream(side)
                    return self.iterate(dt, adaptive)
                ev_a = torch.cuda.Event()
                ev_a.record(main)
                # second stream: pull the ghosts of this stage (the peers' previous stage is complete: barrier below /
                # publish()), then the partition-boundary chunks; then announce this stage once both passes are done
                with torch.cuda.stream(side):
                    P.pull(vin, T[sin])
                    P.stage_part(stage, 2, vin, pv, vout, vol, dt, speed_max=sm, dt_dev=dtd)
                    self.ev_b = torch.cuda.Event()
                    self.ev_b.record(side)
                    side.wait_event(ev_a)
                    if stage == 3:
                        m.barrier(self.speed_loc, self.speed_max)     # stage barrier + CFL max over the ranks
                        self.speed_loc.zero_()                        # both passes of the next stage 3 max into it
                        if adaptive:
                            self.tb.timestep(self.speed_max, CFL, self.length, dt, self.dt_dev)
                            self.ev_b = torch.cuda.Event()            # the next step's kernels read the new dt
                            self.ev_b.record(side)
                    else:
                        m.barrier()
            self.side_done = torch.cuda.Event()
            self.side_done.record(side)
            return 15 + int(adaptive)                     # 3 x (interior pass, pull, boundary pass, barrier) + memsets
        if self.sync == "pull":
            m, P = self.mail, self.plan
            P.pull(prev, T[self.prev])                    # the peers' U^n is complete (barrier of the previous step)
            P.stage(1, prev, None, s1, vol, dt, dt_dev=dtd)
            m.barrier()
            P.pull(s1, T[1])
            P.stage(2, s1, prev, s2, vol, dt, dt_dev=dtd)
            m.barrier()
            P.pull(s2, T[2])
            P.stage(3, s2, prev, nxt, vol, dt, speed_max=self.speed_loc, dt_dev=dtd)
            m.barrier(self.speed_loc, self.speed_max)     # stage barrier + CFL max over the ranks
            if adaptive:
                self.tb.timestep(self.speed_max, CFL, self.length, dt, self.dt_dev)
            return 9 + int(adaptive)                      # 3 x (pull, stage, barrier)
        if self.sync == "kernel":
            m = self.mail
            self.plan.stage(1, prev, None, s1, vol, dt, in_all=T[self.prev], dt_dev=dtd, sync=m)
            self.plan.stage(2, s1, prev, s2, vol, dt, in_all=T[1], dt_dev=dtd, sync=m)
            main = torch.cuda.current_stream()
            if self.side_done is not None:
                main.wait_event(self.side_done)           # the previous reduction has read speed_loc
                self.side_done = None
            self.plan.stage(3, s2, prev, nxt, vol, dt, in_all=T[2], speed_max=self.speed_loc, dt_dev=dtd, sync=m)
            if adaptive:                                  # the next step's kernels need the reduced value
                m.barrier(self.speed_loc, self.speed_max)
                self.tb.timestep(self.speed_max, CFL, self.length, dt, self.dt_dev)
                return 5
            if os.environ.get("T8B200_CFL_SIDE", "1") == "0":   # experiment: the reduction on the compute stream
                m.barrier(self.speed_loc, self.speed_max)
                return 4
            ev = torch.cuda.Event()
            ev.record(main)
            self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                m.barrier(self.speed_loc, self.speed_max)
                self.side_done = torch.cuda.Event()
                self.side_done.record(self.side)
            return 4                                               # 3 stage kernels + the CFL reduction
        if self.mail is not None:   # "peer": barriers and the CFL reduction through the peers' mailboxes
            bar = (lambda *a: None) if self.sync == "none" else self.mail.barrier
            self.plan.stage(1, prev, None, s1, vol, dt, in_all=T[self.prev], dt_dev=dtd)
            bar()
            self.plan.stage(2, s1, prev, s2, vol, dt, in_all=T[1], dt_dev=dtd)
            bar()
            self.plan.stage(3, s2, prev, nxt, vol, dt, in_all=T[2], speed_max=self.speed_loc, dt_dev=dtd)
            self.mail.barrier(self.speed_loc, self.speed_max)
            if adaptive:
                self.tb.timestep(self.speed_max, CFL, self.length, dt, self.dt_dev)
            return 6                                               # 3 stage kernels + 3 barrier kernels
        self.plan.stage(1, prev, None, s1, vol, dt, in_all=T[self.prev], dt_dev=dtd)
        stage_barrier(dist, self.token)                            # device-side, on the compute stream
        self.plan.stage(2, s1, prev, s2, vol, dt, in_all=T[1], dt_dev=dtd)
        stage_barrier(dist, self.token)
        self.plan.stage(3, s2, prev, nxt, vol, dt, in_all=T[2], speed_max=self.speed_max, dt_dev=dtd)
        global_max_wave_speed(dist, self.speed_max)                # barrier + global CFL reduction
        if adaptive:
            self.tb.timestep(self.speed_max, CFL, self.length, dt, self.dt_dev)
        return 3

    def breakdown(self, dt, steps=5):
        """Where a multi-GPU step goes (push mode): CUDA events around every launch of `steps` steps, per rank.  Stands
        in for a profiler capture of a multi-process run (ncu on one rank of a torchrun job hangs in the IPC set-up):
        stage kernels, ghost pushes (with the NVLink bytes they move) and barrier kernels (launch + wait for the
        slowest rank), averaged per launch."""
        if self.sync not in ("push", "fpush", "pushb"):
            return None
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
        marks, st = [], torch.cuda.current_stream()
        for _ in range(steps):
            self.next, self.prev = self.prev, self.next
            prev, s1, s2, nxt = (self.variables(s) for s in (self.prev, 1, 2, self.next))
            vol, T = self.volume(), self.tables
            for stage, vin, vout, so in ((1, prev, s1, 1), (2, s1, s2, 2), (3, s2, nxt, self.next)):
                e = [ev() for _ in range(4)]
                e[0].record(st)
                self.plan.stage(stage, vin, prev if stage > 1 else None, vout, vol, dt,
                                speed_max=self.speed_loc if stage == 3 else None)
                e[1].record(st)
                self.tb.ghost_push(*self.send, vout, T[so])
                e[2].record(st)
                if stage == 3:
                    self.mail.barrier(self.speed_loc, self.speed_max)
                else:
                    self.mail.barrier()
                e[3].record(st)
                marks.append(e)
        torch.cuda.synchronize()
        n = len(marks)
        t = [sum(m[i].elapsed_time(m[i + 1]) for m in marks) / n * 1e3 for i in range(3)]
        n_send = int(self.send[0].numel())
        esz = 8 if self.dtype == torch.float64 else 4
        return {"stage_kernel_us": round(t[0], 1), "ghost_push_us": round(t[1], 1), "barrier_us": round(t[2], 1),
                "ghost_push_bytes": n_send * 5 * esz, "ghost_push_gbs": round(n_send * 5 * esz / (t[1] * 1e-6) / 1e9, 1),
                "elements_pushed": n_send, "note": "per launch, this rank, events between the launches (they serialise "
                "the stream a little: the sum is above the timed ms_per_step / 3)"}

    def drain(self):
        """Joins the side stream (the last CFL reduction) into the compute stream."""
        if self.side_done is not None:
            torch.cuda.current_stream().wait_event(self.side_done)
            self.side_done = None

    def close(self):
        self.drain()
        torch.cuda.synchronize()
        dist.barrier()
        self.plan = None
        if self.mail is not None:
            self.mail.close()
        self.shared.close()


def bitwise_parity(dtype, rank, world, device, level=4, steps=3):
    """Owner-computes is deterministic: the N-rank run must be BITWISE equal to the one-rank run of the same brick.
    Every rank steps the small brick under the production protocol (self-ordering kernels, peer ghost reads), then
    compares its slice with a one-rank solve of the whole brick done on its own GPU."""
    import t8gpu_b200 as tb
    dt = 0.1 * 2.0 ** -level
    sol = MultiGpuEuler(level, dtype, rank, world, device)
    c1 = tb.cartesian_uniform_connectivity(3, level, dtype, 1, 0, device=device, brick=BRICK[world])
    one = tb.EulerSolver(tb.conn_to_host(c1), c1["volumes"].cpu().numpy(), dtype, device=device)
    tb.init_kelvin_helmholtz(3, c1["centroids"], one.variables(one.next))
    n1 = int(c1["n_local"])
    off = [n1 * r // world for r in range(world + 1)]
    equal_ic = bool(torch.equal(one.state()[:, off[rank]:off[rank + 1]], sol.state()))
    for _ in range(steps):
        one.iterate(dt)
        sol.iterate(dt)
    sol.drain()
    torch.cuda.synchronize()
    same = equal_ic and bool(torch.equal(one.state()[:, off[rank]:off[rank + 1]], sol.state()))
    vm = abs(float(one.max_wave_speed().item()) - float(sol.speed_max.item())) == 0.0
    flag = torch.tensor([1.0 if (same and vm) else 0.0], dtype=torch.float64, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    sol.close()
    ok = bool(flag.item() == 1.0)
    return {"vs": "one-rank run of the same brick (level %d trees, %d elements)" % (level, n1), "steps": steps,
            "bitwise_equal": ok, "checked": "conserved variables of every rank + reduced max wave speed"}


def run_multi(args, rank, world, device):
    from bench import ALG_BYTES, ClockSampler, measured_peak
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    ncpu, numa_node, mempolicy = pin_to_gpu_numa_node(device)
    numa_all = [None] * world
    dist.all_gather_object(numa_all, (ncpu, numa_node, mempolicy))
    parity = bitwise_parity(dtype, rank, world, device)
    if os.environ.get("T8B200_SYNC") != "none":   # (the timing experiment without stage ordering is expected to differ)
        assert parity["bitwise_equal"], "the %d-rank run differs from the one-rank run of the same brick" % world
    t0 = time.time()
    sol = MultiGpuEuler(args.level, dtype, rank, world, device)
    t_setup = time.time() - t0
    dt = 0.1 * 2.0 ** -args.level
    n = sol.n
    stream = torch.cuda.current_stream()
    # settle: NVLink links and peer mappings of a box that sat idle come up during the first few hundred milliseconds of
    # traffic (measured: the first run on a fresh box 2.60 ms per step, every later one 2.46); untimed, before the W
    # warm-up steps, after which the state is the KH initial state again
    settle = int(os.environ.get("T8B200_SETTLE_STEPS", "120"))
    u_init = sol.state().clone() if settle > 0 else None
    for _ in range(settle):
        sol.iterate(dt)
    if settle > 0:
        sol.drain()
        torch.cuda.synchronize()
        dist.barrier()
        sol.state().copy_(u_init)
        del u_init
        sol.publish()
    for _ in range(args.warmup):
        sol.iterate(dt)
    sol.drain()
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
    sol.drain()
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    tw1 = time.time()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    every = [torch.zeros_like(ms) for _ in range(world)]
    dist.all_gather(every, ms)
    per_rank_ms = [round(float(t.item()) / args.steps, 4) for t in every]
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    vmax = float(sol.speed_max.item())
    assert vmax > 0 and vmax == vmax

    # e2e: the same job through host buffers: state uploaded from pinned host memory, K x [iterate with the CFL rule:
    # global max wave speed -> next dt, both kept on the device; dt and vmax of every step are copied to the host
    # asynchronously], state downloaded.  No host synchronisation inside the loop; max over ranks.
    u_host = torch.empty((5, n), dtype=dtype).pin_memory()
    u_host.copy_(sol.state())
    out_host = torch.empty((5, n), dtype=dtype).pin_memory()
    hist = torch.zeros((args.steps, 2), dtype=dtype).pin_memory()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sol.state().copy_(u_host, non_blocking=True)
    sol.dt_dev.fill_(dt)
    sol.publish()                    # every rank's upload is complete before neighbours read it
    if sol.sync == "nccl":
        dist.all_reduce(sol.token, op=dist.ReduceOp.MAX)
    for k in range(args.steps):
        sol.iterate(dt, adaptive=True)
        hist[k, 0:1].copy_(sol.dt_dev, non_blocking=True)
        hist[k, 1:2].copy_(sol.speed_max, non_blocking=True)
    out_host.copy_(sol.state(), non_blocking=True)
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    assert bool((hist[:, 0] > 0).all()) and bool((hist[:, 0] <= dt).all()) and bool((hist[:, 1] > 0).all())

    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    breakdown = sol.breakdown(dt)
    info = sol.plan.info
    nfaces, nghost, sync, t_plan = sol.n_faces, sol.n_ghost, sol.sync, sol.t_plan
    sol.close()
    secondary = {}
    if not getattr(args, "no_secondary", False):
        # BASELINE config 3 across the same ranks: adapt + repartition every 10 steps (short run, compared with the
        # one-rank run of the same forest sequence)
        from bench_amr import amr_secondary
        secondary["amr_c3"] = amr_secondary(args.dtype, rank, world, device, dist)
        secondary["subgrid_amr_c4"] = amr_secondary(args.dtype, rank, world, device, dist, subgrid=True)
        # BASELINE config 5 (strong scaling: the SAME mixed-element mesh on every GPU count)
        import argparse
        from bench_hybrid import run_hybrid
        h = run_hybrid(argparse.Namespace(dtype=args.dtype, level=24, steps=20, warmup=3), rank, world, device, dist,
                       emit=False)
        secondary["hybrid_c5"] = {"workload": h["config"]["workload"], "ms_per_step": h["ms_per_step"],
                                  "value": h["value"], "scaling": "strong", "roofline_frac_per_gpu": h["roofline"]["frac"],
                                  "parity": h["parity"]}
    if rank == 0:
        esz = 8 if args.dtype == "f64" else 4
        total = n * world
        peak, src = measured_peak()
        alg = ALG_BYTES[("hex", args.dtype)]
        ms_per_step = ms / args.steps
        achieved = alg * n / (ms_per_step * 1e-3) / 1e9   # per GPU
        state_bytes = 5 * n * esz
        sync_text = {"fpush": "ghost tail filled by the stage kernel: the thread that writes an element its peers hold a "
                              "copy of also stores the new values into those copies over NVLink (posted stores from the "
                              "epilogue); one mailbox barrier kernel per stage publishes them; the stage-3 barrier "
                              "carries the CFL max",
                     "pushb": "ghost tail; after each stage ONE kernel pushes the elements the peers hold copies of into "
                              "their tails over NVLink (consecutive remote stores) and, from its last CTA, runs the "
                              "mailbox barrier that publishes them; the single-rank stage kernel follows (no peer memory "
                              "inside it); the stage-3 barrier carries the CFL max",
                     "push": "ghost tail; after each stage every rank pushes the elements its peers hold copies of into "
                             "their tails over NVLink (consecutive remote stores), one mailbox barrier kernel publishes "
                             "them, the single-rank stage kernel follows (no peer memory inside it); the stage-3 barrier "
                             "carries the CFL max",
                     "overlap": "ghost tail; every stage in two passes: the chunks without ghosts on the compute stream right "
                                "behind the previous stage, on a second stream mailbox barrier -> pull kernel (ghost "
                                "copies from the peers' rows over NVLink) -> the partition-boundary chunks; the stage-3 "
                                "barrier carries the CFL max",
                     "pull": "ghost tail: per stage one peer-memory mailbox barrier kernel + one pull kernel copying the "
                             "ghosts from the peers' rows over NVLink into local copies behind the own elements, then "
                             "the single-rank stage kernel (no peer memory inside it); the stage-3 barrier carries the "
                             "CFL max",
                     "kernel": "stage kernels order themselves through peer-memory mailboxes (ghost-reading chunks first: "
                               "wait for the peers' previous stage, last one signals; interior chunks never wait); CFL "
                               "max over ranks = one 32-thread mailbox kernel per step",
                     "peer": "peer-memory mailbox barrier kernel per RK stage on the compute stream; stage 3 carries "
                             "the max wave speed",
                     "nccl": "1 NCCL all-reduce (1 element) per RK stage on the compute stream; stage 3 carries the max "
                             "wave speed",
                     "none": "NO stage ordering (timing experiment, results invalid)"}[sync]
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
                           "sync": sync_text, "l2": "inputs larger than L2", "settle_steps_before_warmup": settle,
                           "host_setup_s": round(t_setup, 2), "tile_plan_s": round(t_plan, 3),
                           "tile_plan_built_on": info.get("built_on", "host"), "host_cores": os.cpu_count(),
                           "host_cpus_bound_to_gpu_numa_node": ncpu,
                           "per_rank_(local_cpus, numa_node, mempolicy_preferred_set)": numa_all,
                           "per_rank_ms_per_step": per_rank_ms,
                           "step_breakdown_rank0": breakdown,
                           "plan": info},
                "clocks": clocks, "parity": parity, "secondary": secondary,
                "e2e": {"value": total * args.steps / (e2e_ms * 1e-3), "unit": "cell-updates/s",
                        "h2d_bytes_per_step": (state_bytes / args.steps + esz) * world,
                        "d2h_bytes_per_step": (state_bytes / args.steps + 2 * esz) * world,
                        "ms_per_step": e2e_ms / args.steps,
                        "protocol": "pinned-host state in, K x (iterate with dt on the device: global CFL max -> next "
                                    "dt, no host synchronisation; async D2H of dt and vmax per step), state out"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": src,
                             "kernel": "structured_stage_kernel", "note": "per GPU"},
                "max_wave_speed": vmax}
        print(json.dumps(line))
    dist.destroy_process_group()
