#!/usr/bin/env python
"""AMR cycle of BASELINE config 3 on one B200 (measurement helper for DESIGN.md, run by hand under gpurun; lives in
tests/ because the mini-forest of the oracle stands in for t8code, which owns adapt / balance on the host).

Every `--every` steps: gradient criterion (GPU) -> criteria to the host -> forest adapt + balance (host, mini-forest)
-> old->new index map -> device remap of the state -> connectivity (host) -> tile plan rebuild.  Reports the pure
stepping rate between adapts and the end-to-end rate with the cycle broken out.

    python tests/perf_amr.py [--level 6] [--cycles 4] [--every 10] [--dtype f64]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--max-level", type=int, default=8)
    ap.add_argument("--cycles", type=int, default=4)
    ap.add_argument("--every", type=int, default=10)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--subgrid", action="store_true",
                    help="config 4: Subgrid<4,4,4>, the subgrid criterion (threshold 0.02), cell-level remap")
    args = ap.parse_args()
    import torch
    import oracle
    import t8gpu_b200 as tb
    npdt = np.float64 if args.dtype == "f64" else np.float32
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    dev = torch.device("cuda", 0)
    f = oracle.Forest(3, args.level)
    lv, cent, vol, _ = f.elements()
    S = 64 if args.subgrid else 1

    def make_solver(forest, volumes):
        if args.subgrid:
            return tb.SubgridEulerSolver(forest.connectivity(subgrid=True, dtype=npdt), volumes.astype(npdt), tdt,
                                         device=dev, mode="fused")
        return tb.EulerSolver(forest.connectivity(dtype=npdt), volumes.astype(npdt), tdt, device=dev, mode="fused",
                              max_level=args.max_level)

    def volume_of(solver):
        return solver.vol if args.subgrid else solver.volume()

    u0 = (oracle.subgrid_init_kh(3, cent.astype(npdt), lv, npdt) if args.subgrid
          else oracle.init_kh_points(3, cent.astype(npdt), npdt))
    sol = make_solver(f, vol)
    sol.set_state(u0)
    dt = 0.1 * 2.0 ** -(args.max_level + (2 if args.subgrid else 0))
    for _ in range(3):
        sol.iterate(dt)
    torch.cuda.synchronize()
    t_step = t_crit = t_forest = t_conn = t_plan = t_remap = 0.0
    updates = 0
    hist = []
    wall0 = time.time()
    for cyc in range(args.cycles):
        n = sol.nc if args.subgrid else sol.n
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.every):
            sol.iterate(dt)
        e1.record()
        torch.cuda.synchronize()
        t_step += e0.elapsed_time(e1) * 1e-3
        updates += n * args.every
        # criteria on the device (threshold of the example: refine where the scaled density jump is large)
        t = time.time()
        if args.subgrid:
            crit_h = tb.subgrid_criteria(3, sol.state()[0], sol.vol).cpu().numpy()
            threshold = 0.02    # examples/subgrid: refine above, coarsen below (subgrid_mesh_manager.inl adapt callback)
        else:
            crit = tb.gradient_criteria(sol.plan, sol.state()[0], sol.volume())
            # the example's threshold (10) is tuned for its shell mesh; scale so that the shear layers refine
            crit_h = (crit * (10.0 / 0.5)).cpu().numpy()
            threshold = 10.0
        torch.cuda.synchronize()
        t_crit += time.time() - t
        t = time.time()
        f2 = f.adapt(crit_h, threshold, 1, args.max_level)
        amap = f.adapt_map(f2)
        t_forest += time.time() - t
        t = time.time()
        lv2, cent2, vol2, _ = f2.elements()
        conn2 = f2.connectivity(subgrid=args.subgrid, dtype=npdt)
        t_conn += time.time() - t
        t = time.time()
        if args.subgrid:
            new = tb.SubgridEulerSolver(conn2, vol2.astype(npdt), tdt, device=dev, mode="fused")
        else:
            new = tb.EulerSolver(conn2, vol2.astype(npdt), tdt, device=dev, mode="fused", max_level=args.max_level)
        torch.cuda.synchronize()
        t_plan += time.time() - t
        t = time.time()
        tb.adapt_remap(torch.as_tensor(amap).to(dev), sol.variables(sol.next), new.variables(new.next), volume_of(sol),
                       volume_of(new), 3 if args.subgrid else 0)
        torch.cuda.synchronize()
        t_remap += time.time() - t
        hist.append(dict(cycle=cyc, cells=int(n), new_cells=int(new.nc if args.subgrid else new.n),
                         chunks=int(new.plan.info["n_chunks"]),
                         max_halo=int(new.plan.info["max_halo"]), max_faces=int(new.plan.info["max_faces"])))
        sol, f = new, f2
        assert torch.isfinite(sol.state()).all()
    wall = time.time() - wall0
    print(json.dumps({
        "dtype": args.dtype, "subgrid": args.subgrid, "every": args.every, "cycles": args.cycles, "history": hist,
        "stepping_cell_updates_per_s": updates / t_step, "end_to_end_cell_updates_per_s": updates / wall,
        "seconds": {"stepping": t_step, "criteria_gpu+d2h": t_crit, "forest_adapt_host(mini-forest)": t_forest,
                    "connectivity_host(mini-forest)": t_conn, "plan_build+alloc": t_plan, "remap_gpu": t_remap,
                    "wall": wall}, "host_cores": os.cpu_count()}))


if __name__ == "__main__":
    main()
