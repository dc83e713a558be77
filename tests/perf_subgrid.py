#!/usr/bin/env python
"""Throughput of the Subgrid<4,4,4> path (BASELINE config 4 per-GPU size: uniform periodic hex level 6 = 262 144
elements x 64 = 16.8 M cells) on one B200: fused stage kernel vs the reference-shaped schedule vs the reference's own
kernels (oracle/_ref).  Measurement helper for DESIGN.md, run by hand under gpurun; lives in tests/ because it takes
mesh and initial data from the oracle.

    python tests/perf_subgrid.py [--level 6] [--steps 20] [--dtype f32|f64] [--amr]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ALG = {"f32": 165.8, "f64": 328.3}   # algorithmic bytes per cell per RK3 step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--amr", action="store_true", help="refine a slab one level (2:1 hanging faces)")
    ap.add_argument("--modes", default="fused,unfused,reference")
    args = ap.parse_args()
    import torch
    import oracle
    import t8gpu_b200
    from oracle import ref_cuda
    npdt = np.float64 if args.dtype == "f64" else np.float32
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    dev = torch.device("cuda", 0)
    f = oracle.Forest(3, args.level)
    if args.amr:
        lv, cent, vol, _ = f.elements()
        f = f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < 0.1, 1.0, 0.0), 0.02, 1, args.level + 1)
    lv, cent, vol, _ = f.elements()
    conn = f.connectivity(subgrid=True, dtype=npdt)
    u0 = oracle.subgrid_init_kh(3, cent.astype(npdt), lv, npdt)
    ncell = u0.shape[1]
    dt = 0.1 * 2.0 ** -(args.level + 3)
    peak = 6551.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"cells": ncell, "elements": int(conn["n_local"]), "dtype": args.dtype, "amr": args.amr}
    for mode in args.modes.split(","):
        if mode == "reference":
            if args.amr or not ref_cuda.available():
                continue
            s = ref_cuda.RefSolver("sg", npdt, 3, args.level, True)
            s.set_state(u0)
            ms = s.time_steps(dt, args.warmup, args.steps)
            s.close()
        else:
            sol = t8gpu_b200.SubgridEulerSolver(conn, vol.astype(npdt), tdt, device=dev, mode=mode)
            sol.set_state(u0)
            for _ in range(args.warmup):
                sol.iterate(dt)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                sol.iterate(dt)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            assert torch.isfinite(sol.state()).all()
            del sol
        rate = ncell * args.steps / (ms * 1e-3)
        out[mode] = {"ms_per_step": ms / args.steps, "cell_updates_per_s": rate,
                     "roofline_frac": ALG[args.dtype] * rate / 1e9 / peak}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
