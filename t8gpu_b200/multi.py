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


def adapt_partition_ranges(amap, off_old, off_new, rank):
    """Index arithmetic of one adapt + repartition cycle for `rank` (host side; the forest work itself is t8code's).

    amap: the global old -> new element map of the adapted forest (n_new + 1 entries, mesh_manager.inl:258-281);
    off_old / off_new: partition offsets before the adapt / after the repartition (nranks + 1 entries).
    t8code adapts every rank's elements in place (a family is only coarsened inside one rank), so the adapted, not yet
    repartitioned forest keeps rank q's elements in [lo[q], lo[q+1]).  Returns
      lo          those ranges (nranks + 1),
      adapt_data  this rank's local map for t8b200_adapt_remap (n_mid + 1 entries, relative to its old elements),
      owner, index  for every element of this rank's NEW partition: the rank that holds it after the adapt and its
                  index there (what t8_forest_partition_data delivers, mesh_manager.inl:655-676) for
                  t8b200_partition_remap."""
    import numpy as np
    amap = np.asarray(amap, np.int64)
    off_old, off_new = np.asarray(off_old, np.int64), np.asarray(off_new, np.int64)
    n_new = len(amap) - 1
    lo = np.searchsorted(amap[:-1], off_old, side="left").astype(np.int64)
    lo[-1] = n_new
    adapt_data = (amap[lo[rank]:lo[rank + 1] + 1] - off_old[rank]).astype(np.int32)
    g = np.arange(off_new[rank], off_new[rank + 1])
    owner = (np.searchsorted(lo, g, side="right") - 1).astype(np.int32)
    index = (g - lo[owner]).astype(np.int32)
    return lo, adapt_data, owner, index


def send_csr(send, n_local, device):
    """The send list (src_idx, dst_rank, dst_idx) of send_lists grouped by source element: (send_off[n_local + 1],
    send_rank, send_idx) device int32 tensors for t8b200_fused_stage_push_*."""
    import numpy as np
    import torch
    src, drk, dix = (t.cpu().numpy() for t in send)
    order = np.argsort(src, kind="stable")
    off = np.zeros(n_local + 1, np.int64)
    np.add.at(off, src[order].astype(np.int64) + 1, 1)
    off = np.cumsum(off).astype(np.int32)
    mk = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(device)  # noqa: E731
    if len(src) == 0:
        return mk(off), mk(np.zeros(1)), mk(np.zeros(1))
    return mk(off), mk(drk[order]), mk(dix[order])
