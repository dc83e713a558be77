"""Host-side logic of the one-process-per-GPU path (harness glue for tests and bench.py, on top of the C ABI).

Replaces what SharedDeviceVector does with MPI in the reference (t8gpu/memory/shared_device_vector.inl:171-198): every
rank allocates its state buffer with t8b200_shared_alloc, the 64-byte cudaIpc handles (+ row capacity) travel through a
`torch.distributed` all-gather (NCCL on the GPUs, gloo in the CPU tests), peers are mapped with t8b200_shared_open and
the `[var][rank] -> pointer` tables are filled with row addresses.  The stage ordering and the CFL reduction are
all-reduces on the compute stream (solver.cu:98-99, :219-223 use cudaDeviceSynchronize + MPI_Barrier / MPI_Allreduce).

Nothing here touches the GPU except through the tensors it is handed, so the same functions run under gloo on CPU.
"""
WIRE_BYTES = 72   # 64-byte IPC handle + 8-byte row capacity (elements)

BRICK = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}   # weak-scaling meshes: one tree per GPU


def pack_wire(handle, capacity):
    """(64-byte handle, capacity) -> list of WIRE_BYTES ints."""
    handle = bytes(handle)
    if len(handle) != 64:
        raise ValueError("a cudaIpcMemHandle_t is 64 bytes, got %d" % len(handle))
    if capacity < 0:
        raise ValueError("negative capacity")
    return list(handle) + list(int(capacity).to_bytes(8, "little"))


def unpack_wire(raw):
    raw = bytes(int(x) for x in raw)
    if len(raw) != WIRE_BYTES:
        raise ValueError("wire record of %d bytes" % len(raw))
    return raw[:64], int.from_bytes(raw[64:72], "little")


def exchange_wires(dist, handle, capacity, world, device):
    """All-gather of every rank's (handle, capacity); returns a list over ranks."""
    import torch
    mine = torch.tensor(pack_wire(handle, capacity), dtype=torch.uint8, device=device)
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    return [unpack_wire(t.cpu().tolist()) for t in every]


def row_pointers(bases, caps, nvar, nsteps, esz):
    """[step][var][rank] -> address of that row in rank's buffer (MemoryManager layout: row step*nvar+var, stride = that
    rank's capacity), from the mapped base pointers."""
    if len(bases) != len(caps):
        raise ValueError("one base pointer and one capacity per rank")
    return [[[int(bases[r]) + (s * nvar + k) * int(caps[r]) * esz for r in range(len(bases))]
             for k in range(nvar)] for s in range(nsteps)]


def stage_barrier(dist, token):
    """Orders an RK stage across ranks on the stream the collective is enqueued on (no host synchronisation)."""
    dist.all_reduce(token, op=dist.ReduceOp.MAX)


def global_max_wave_speed(dist, speed_max):
    """MPI_Allreduce(MAX) of solver.cu:219-223; doubles as the stage-3 barrier."""
    dist.all_reduce(speed_max, op=dist.ReduceOp.MAX)
    return speed_max


def timestep(vmax, cfl, max_level, dt_cap=None):
    """CompressibleEulerSolver::compute_timestep (solver.cu:225-228): cfl * 0.5^max_level / vmax, optionally capped."""
    dt = cfl * 0.5 ** max_level / vmax
    return dt if dt_cap is None else min(dt_cap, dt)


def send_lists(dist, plan, n_local, rank, world, device):
    """Push lists of this rank from the pull lists of all ranks: what every peer p pulls from `rank` (p's plan arrays 17 /
    18), with its destination n_local(p) + j in p's rows.  Returns device int32 tensors (src_idx, dst_rank, dst_idx),
    sorted by destination.  One all-gather of the padded pull lists at set-up."""
    import numpy as np
    import torch
    pr, pi = plan.device_array(17), plan.device_array(18)
    meta = torch.tensor([len(pr), n_local], dtype=torch.int64, device=device)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [m.cpu().tolist() for m in metas]
    cap = max(1, max(m[0] for m in metas))
    mine = torch.zeros((2, cap), dtype=torch.int32, device=device)
    mine[0, :len(pr)] = torch.as_tensor(pr)
    mine[1, :len(pi)] = torch.as_tensor(pi)
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    src, drk, dix = [], [], []
    for p in range(world):
        if p == rank:
            continue
        n, nl = metas[p]
        rk, ix = every[p][0, :n].cpu().numpy(), every[p][1, :n].cpu().numpy()
        sel = np.nonzero(rk == rank)[0]
        src.append(ix[sel])
        drk.append(np.full(len(sel), p, np.int32))
        dix.append((nl + sel).astype(np.int32))
    cat = lambda a: torch.as_tensor(np.concatenate(a) if a else np.zeros(0, np.int32)).to(torch.int32).to(device)  # noqa: E731
    return cat(src), cat(drk), cat(dix)
