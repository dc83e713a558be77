#!/usr/bin/env python
"""Benchmark of the t8gpu hot path on B200: cell-updates/s per SSP-RK3 step and fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Kelvin-Helmholtz on a uniform periodic 3-D hex forest, level 8 = 16 777 216
elements per GPU, fp64, fixed dt = 0.1 * 2^-8, no adaptation.  One "step" = one iterate() = 3 RK stages, each a flux
evaluation over all faces + the stage update.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic (compulsory) bytes per element per RK3 step, SURVEY.md 8(d) / DESIGN.md "Roofline"
ALG_BYTES = {("hex", "f64"): 704, ("hex", "f32"): 388, ("quad", "f32"): 316, ("quad", "f64"): 560}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.2] or [l for (_, l) in self.lines]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_problem(level, dtype_name, rank, nranks, device, rotate=False, mode="fused"):
    """Device connectivity + KH state for this rank's partition; returns (solver, info)."""
    import torch
    import t8gpu_b200 as tb
    dtype = torch.float64 if dtype_name == "f64" else torch.float32
    t0 = time.time()
    conn = tb.cartesian_uniform_connectivity(3, level, dtype, nranks, rank, device=device)
    torch.cuda.synchronize()
    t_conn = time.time() - t0
    n = int(conn["n_local"])
    t0 = time.time()
    host = tb.conn_to_host(conn)
    if rotate:
        # same topology, the whole mesh rotated in space: general unit normals -> the uncompressed-geometry path of
        # the tile plan (what a non-Cartesian / mixed-element mesh takes)
        import numpy as np
        q, _ = np.linalg.qr(np.random.default_rng(7).normal(size=(3, 3)))
        nrm = host["face_normals"].reshape(-1, 3).astype(np.float64) @ q.T
        host["face_normals"] = np.ascontiguousarray(nrm.reshape(-1).astype(host["face_normals"].dtype))
    sol = tb.EulerSolver(host, host["volumes"], dtype, device=device, mode=mode)
    torch.cuda.synchronize()
    t_plan = time.time() - t0
    tb.init_kelvin_helmholtz(3, conn["centroids"], sol.variables(sol.next))
    torch.cuda.synchronize()
    info = dict(n=n, faces=int(conn["n_faces"]), t_connectivity_s=round(t_conn, 3), t_plan_s=round(t_plan, 3),
                plan=sol.plan.info if sol.plan is not None else None)
    del conn
    return sol, info


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.workload == "subgrid":
        from bench_subgrid import run_subgrid
        run_subgrid(args, rank, world, device)
        if world > 1:
            dist.destroy_process_group()
        return
    if world > 1:
        from bench_multigpu import run_multi
        return run_multi(args, rank, world, device)

    dtype_name = args.dtype
    dt = 0.1 * 2.0 ** -args.level
    sol, info = build_problem(args.level, dtype_name, 0, 1, device, rotate=args.rotate, mode=args.mode)
    n = info["n"]
    stream = torch.cuda.current_stream()

    # ---------------- device-resident throughput ("value")
    for _ in range(args.warmup):
        sol.iterate(dt)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    torch.cuda.synchronize()
    t0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        launches += sol.iterate(dt)
    ev1.record(stream)
    torch.cuda.synchronize()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    ms = ev0.elapsed_time(ev1)
    ms_per_step = ms / args.steps
    value = n * args.steps / (ms * 1e-3)
    vmax = float(sol.max_wave_speed().item())
    assert vmax > 0 and vmax == vmax, "wave speed is not finite: the run diverged"

    # ---------------- end to end through the public API with host buffers ("e2e")
    # job = upload the initial state from pinned host memory, K x [iterate(dt); read back the stage-3 maximum wave
    # speed (the CFL reduction) and compute the next dt on the host], download the final state.  All inside the
    # timed region; copies are amortised over K steps exactly as in a real run of the reference's main loop.
    u_host = torch.empty((5, n), dtype=sol.dtype).pin_memory()
    u_host.copy_(sol.state())
    out_host = torch.empty((5, n), dtype=sol.dtype).pin_memory()
    vmax_host = torch.empty(1, dtype=sol.dtype).pin_memory()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sol.state().copy_(u_host, non_blocking=True)
    cur_dt = dt
    for _ in range(args.steps):
        sol.iterate(cur_dt)
        vmax_host.copy_(sol.max_wave_speed(), non_blocking=True)
        stream.synchronize()
        # CompressibleEulerSolver::compute_timestep (solver.cu:225-228), capped by the fixed dt of the config
        cur_dt = min(dt, 0.7 * 0.5 ** args.level / float(vmax_host[0]))
    out_host.copy_(sol.state(), non_blocking=True)
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    esz = 8 if dtype_name == "f64" else 4
    state_bytes = 5 * n * esz
    e2e = {"value": n * args.steps / (e2e_ms * 1e-3), "unit": "cell-updates/s",
           "h2d_bytes_per_step": state_bytes / args.steps + esz, "d2h_bytes_per_step": state_bytes / args.steps + esz,
           "ms_per_step": e2e_ms / args.steps,
           "protocol": "pinned-host state in, K x (iterate + D2H max wave speed + host dt), state out"}

    peak, peak_src = measured_peak()
    alg = ALG_BYTES[("hex", dtype_name)]
    achieved = alg * n / (ms_per_step * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % dtype_name)
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic if args.mode == "fused" and not args.rotate else None, "peak_source": peak_src,
                "kernel": "fused_stage_kernel" if args.mode == "fused" else "flux_faces_kernel + rk3_stage_kernel",
                "alg_bytes_per_launch": alg * n / 3.0, "avg_launch_ms": ms_per_step / 3.0}

    line = {"metric": "cell-updates/s per RK3 step", "value": value, "unit": "cell-updates/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_name, "data": "synthetic",
            "config": {"workload": "kelvin_helmholtz 3D uniform periodic hex mesh level %d (%d elements) %s, "
                                   "fixed dt, no adaptation" % (args.level, n, dtype_name),
                       "elements_per_gpu": n, "faces_per_gpu": info["faces"], "l2": "inputs larger than L2 "
                       "(%.0f MB of state per stage)" % (2 * state_bytes / 1e6),
                       "mode": ("fused tile plan" if args.mode == "fused" else "reference-shaped kernels, reference schedule") +
                               (", general normals (mesh rotated)" if args.rotate else ""),
                       "host_setup_s": {"connectivity_device": info["t_connectivity_s"],
                                        "tile_plan_host": info["t_plan_s"], "host_cores": os.cpu_count()}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "max_wave_speed": vmax}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(dtype_name)
    print(json.dumps(line))


def cpu_baseline(dtype_name, budget_s=12.0, level=6):
    """The CPU oracle (kind "port": the reference has no CPU implementation of this path) on a bounded sample of the
    same workload: uniform periodic hex level 6 (262 144 elements), same IC and dt rule, single thread."""
    import numpy as np
    import oracle
    npdt = np.float64 if dtype_name == "f64" else np.float32
    f = oracle.Forest(3, level)
    conn = f.connectivity(dtype=npdt)
    lv, cent, vol, _ = f.elements()
    u = oracle.init_kh_points(3, cent.astype(npdt), npdt)
    vol = vol.astype(npdt)
    dt = 0.1 * 2.0 ** -level
    steps, t0 = 0, time.time()
    while steps < 2 or time.time() - t0 < budget_s:
        u, _, _ = oracle.iterate(conn, vol, u, dt)
        steps += 1
    el = time.time() - t0
    return {"value": f.num_elements * steps / el, "unit": "cell-updates/s", "cores": 1, "kind": "port",
            "sample": "uniform periodic hex level %d (%d elements), %d RK3 steps, %.1f s, %s" %
                      (level, f.num_elements, steps, el, dtype_name)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bench_reference import run_reference_arm
    run_reference_arm(args)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="unstructured", choices=["unstructured", "subgrid"],
                    help="unstructured = BASELINE configs[1] (the headline); subgrid = configs[3], Subgrid<4,4,4>")
    ap.add_argument("--level", type=int, default=None,
                    help="uniform refinement level per GPU (default 8 = 16.8M hexes; subgrid: 6 = 16.8M cells)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rotate", action="store_true", help="rotate the mesh in space: general-normal path of the plan")
    ap.add_argument("--mode", default="fused", choices=["fused", "unfused"],
                    help="unfused = the reference's schedule through the reference-shaped drop-in kernels")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.level is None:
        args.level = 6 if args.workload == "subgrid" else 8
    if args.impl == "reference":
        return run_reference(args)
    import t8gpu_b200
    t8gpu_b200.lib()  # fail loudly if the CUDA extension is missing
    run_ours(args)


if __name__ == "__main__":
    main()
