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
    plan = None
    if mode == "fused" and not rotate:
        # the plan built on the device from the device-resident connectivity: no D2H copy, no host loop
        plan = tb.Plan.from_device(conn, dtype)
    if plan is not None:
        sol = tb.EulerSolver(dict(n_local=n, n_faces=int(conn["n_faces"]), n_bfaces=0), conn["volumes"], dtype,
                             device=device, mode=mode, plan=plan)
    else:
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
                plan=sol.plan.info if sol.plan is not None else None,
                plan_built_on=(sol.plan.info.get("built_on", "host") if sol.plan is not None else None))
    del conn
    return sol, info


def fp64_peak():
    """Measured FP64 rate of this pool's B200 (tools/fp64_peak.cu, record in profiles/r2_fp64_peak.txt): FMA/s."""
    p = os.path.join(ROOT, "profiles", "r2_fp64_peak.txt")
    best = 0.0
    if os.path.exists(p):
        for ln in open(p):
            if "TFMA/s" in ln:
                try:
                    best = max(best, float(ln.split(":")[1].split("TFMA/s")[0]))
                except Exception:
                    pass
    return (best or 16.5) * 1e12, ("profiles/r2_fp64_peak.txt" if best else "fallback 16.5 TFMA/s")


def profile_record(dtype_name):
    """Per-launch counters of the dominant kernel from the committed ncu capture (profiles/traffic_*.json)."""
    tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % dtype_name)
    try:
        return json.load(open(tp))
    except Exception:
        return {}


def time_loop(sol, dt, steps, stream):
    import torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ev0.record(stream)
    for _ in range(steps):
        launches += sol.iterate(dt)
    ev1.record(stream)
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1), launches


def e2e_unstructured(sol, dt, level, steps, stream):
    """The same job through the public API with HOST buffers: state uploaded from pinned host memory, K x [iterate with
    the CFL rule: stage-3 max wave speed -> next dt, both kept on the device (t8b200_timestep); dt and vmax of every
    step copied to the host asynchronously], state downloaded -- all inside the timed region, no host synchronisation
    inside the loop (the reference's compute_timestep returns through the host, solver.cu:214-217)."""
    import torch
    n = sol.n
    u_host = torch.empty((5, n), dtype=sol.dtype).pin_memory()
    u_host.copy_(sol.state())
    out_host = torch.empty((5, n), dtype=sol.dtype).pin_memory()
    hist = torch.zeros((steps, 2), dtype=sol.dtype).pin_memory()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sol.state().copy_(u_host, non_blocking=True)
    sol.dt_dev.fill_(dt)
    for k in range(steps):
        sol.iterate(dt, adaptive=True, length=0.5 ** level)
        hist[k, 0:1].copy_(sol.dt_dev, non_blocking=True)
        hist[k, 1:2].copy_(sol.speed_max, non_blocking=True)
    out_host.copy_(sol.state(), non_blocking=True)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    assert bool((hist[:, 0] > 0).all()) and bool((hist[:, 0] <= dt).all()) and bool((hist[:, 1] > 0).all())
    esz = sol.state().element_size()
    sb = 5 * n * esz
    return {"value": n * steps / (ms * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": sb / steps + esz,
            "d2h_bytes_per_step": sb / steps + 2 * esz, "ms_per_step": ms / steps,
            "protocol": "pinned-host state in, K x (iterate with dt on the device: stage-3 CFL max -> next dt, no host "
                        "synchronisation; async D2H of dt and vmax per step), state out"}


def parity_unstructured(sol, level, dtype_name, steps=3):
    """After the timing: the measured configuration against the reference's OWN CUDA kernels (oracle/_ref, the checker)
    from the same Kelvin-Helmholtz state, relative L-infinity after each of `steps` steps (tolerance: north star)."""
    import numpy as np
    import torch
    tol = 1e-12 if dtype_name == "f64" else 1e-5
    npdt = np.float64 if dtype_name == "f64" else np.float32
    dt = 0.1 * 2.0 ** -level
    try:
        from oracle import ref_cuda
        have_ref = ref_cuda.available()
    except Exception:
        have_ref = False
    if not have_ref:
        return {"vs": "unavailable (oracle/_ref not built)", "ok": None}
    import t8gpu_b200 as tb
    t0 = time.time()
    conn = tb.cartesian_uniform_connectivity(3, level, sol.dtype, 1, 0, device=sol.device)
    tb.init_kelvin_helmholtz(3, conn["centroids"], sol.variables(sol.next))
    del conn
    u0 = sol.state().cpu().numpy().astype(npdt)
    ref = ref_cuda.RefSolver("uns", npdt, 3, level, True)   # the reference's mesh manager + kernels, unmodified
    ref.set_state(u0)
    errs = []
    for _ in range(steps):
        ref.iterate(dt)
        sol.iterate(dt)
        a, b = sol.state().cpu().numpy().astype(np.float64), ref.get_state().astype(np.float64)
        scale = np.abs(b).max(axis=1)
        scale = np.where(scale < 1e-3 * scale.max(), scale.max(), scale)   # tests/util.py: rel_linf
        errs.append(float((np.abs(a - b).max(axis=1) / scale).max()))
    ref.close()
    ok = all(e <= (k + 1) * tol for k, e in enumerate(errs))
    return {"vs": "reference CUDA kernels (oracle/_ref, examples/compressible_euler compiled unmodified)",
            "level": level, "elements": int(u0.shape[1]), "steps": steps, "rel_linf_after_step": errs,
            "tolerance_per_step": tol, "ok": ok, "seconds": round(time.time() - t0, 1),
            "note": "a variable whose magnitude is < 1e-3 of the state is measured against the state scale "
                    "(rho_v2 == 0 in this set-up)"}


def measure_unstructured(level, dtype_name, steps, warmup, device, rotate=False, mode="fused", e2e=True, parity=True,
                         sampler_index=0):
    """One unstructured measurement: device-resident rate, e2e, roofline (HBM + FP64 pipe), parity."""
    import torch
    dt = 0.1 * 2.0 ** -level
    sol, info = build_problem(level, dtype_name, 0, 1, device, rotate=rotate, mode=mode)
    n = info["n"]
    stream = torch.cuda.current_stream()
    for _ in range(warmup):
        sol.iterate(dt)
    torch.cuda.synchronize()
    sampler = ClockSampler(sampler_index)
    sampler.start()
    time.sleep(0.3)
    torch.cuda.synchronize()
    t0 = time.time()
    ms, launches = time_loop(sol, dt, steps, stream)
    clocks = sampler.stop(t0, time.time())
    ms_per_step = ms / steps
    vmax = float(sol.max_wave_speed().item())
    assert vmax > 0 and vmax == vmax, "wave speed is not finite: the run diverged"
    out = {"n": n, "ms_per_step": ms_per_step, "value": n * steps / (ms * 1e-3), "launches": launches,
           "clocks": clocks, "max_wave_speed": vmax, "info": info}
    if e2e and mode == "fused":
        out["e2e"] = e2e_unstructured(sol, dt, level, steps, stream)
    peak, peak_src = measured_peak()
    alg = ALG_BYTES[("hex", dtype_name)]
    achieved = alg * n / (ms_per_step * 1e-3) / 1e9
    rec = profile_record(dtype_name) if (mode == "fused" and not rotate) else {}
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": rec.get("dram_bytes_per_launch"), "peak_source": peak_src,
            "kernel": rec.get("kernel", "structured_stage_kernel" if mode == "fused" else
                              "flux_faces_kernel + rk3_stage_kernel"),
            "alg_bytes_per_launch": alg * n / 3.0, "avg_launch_ms": ms_per_step / 3.0}
    if dtype_name == "f64" and rec.get("fp64_thread_inst_per_launch"):
        # the second roofline: executed FP64 instructions of the committed ncu capture / live launch time
        # against the measured FP64 rate -- this, not HBM, is the pipe the kernel is closest to
        pk, pk_src = fp64_peak()
        rate = rec["fp64_thread_inst_per_launch"] / (ms_per_step / 3.0 * 1e-3)
        roof["fp64_pipe"] = {"achieved": rate, "peak": pk, "unit": "FP64 thread-instructions/s", "frac": rate / pk,
                             "thread_inst_per_launch": rec["fp64_thread_inst_per_launch"],
                             "per_element_stage": rec["fp64_thread_inst_per_launch"] / n, "peak_source": pk_src,
                             "floor_ms_per_step": 3e3 * rec["fp64_thread_inst_per_launch"] / pk}
        roof["limiter"] = ("latency of the per-face FP64 dependency chain at 24 warps per SM (ncu: FP64 pipe %s %% busy, "
                           "issue slots %s %%, DRAM %s %% of peak): neither roofline is saturated; the FP64 pipe is the "
                           "nearer one" % (rec.get("fp64_pipe_busy_pct", "52-58"), rec.get("issue_active_pct", "54-57"),
                                           rec.get("dram_pct_of_peak", "26-35")))
    out["roofline"] = roof
    if parity and mode == "fused" and not rotate:
        out["parity"] = parity_unstructured(sol, level, dtype_name)
    return out


def short(m, keys=("ms_per_step", "value", "clocks", "parity")):
    d = {k: m[k] for k in keys if k in m}
    d["roofline_frac"] = m["roofline"]["frac"]
    d["roofline_achieved_gbs"] = m["roofline"]["achieved"]
    if "e2e" in m:
        d["e2e_ms_per_step"] = m["e2e"]["ms_per_step"]
    return d


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
    if args.workload == "hybrid":
        from bench_hybrid import run_hybrid
        run_hybrid(args, rank, world, device, dist if world > 1 else None)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "amr":
        from bench_amr import run_amr
        run_amr(args, rank, world, device, dist if world > 1 else None)
        if world > 1:
            dist.destroy_process_group()
        return
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
    m = measure_unstructured(args.level, dtype_name, args.steps, args.warmup, device, rotate=args.rotate,
                             mode=args.mode, parity=not args.no_parity, sampler_index=local)
    n, info = m["n"], m["info"]
    esz = 8 if dtype_name == "f64" else 4
    line = {"metric": "cell-updates/s per RK3 step", "value": m["value"], "unit": "cell-updates/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_name, "data": "synthetic",
            "config": {"workload": "kelvin_helmholtz 3D uniform periodic hex mesh level %d (%d elements) %s, "
                                   "fixed dt, no adaptation" % (args.level, n, dtype_name),
                       "elements_per_gpu": n, "faces_per_gpu": info["faces"], "l2": "inputs larger than L2 "
                       "(%.0f MB of state per stage)" % (2 * 5 * n * esz / 1e6),
                       "mode": ("fused tile plan" if args.mode == "fused" else "reference-shaped kernels, reference schedule") +
                               (", general normals (mesh rotated)" if args.rotate else ""),
                       "host_setup_s": {"connectivity_device": info["t_connectivity_s"],
                                        "tile_plan": info["t_plan_s"], "tile_plan_built_on": info["plan_built_on"],
                                        "host_cores": os.cpu_count()}},
            "clocks": m["clocks"], "gpu_launches": m["launches"], "roofline": m["roofline"],
            "max_wave_speed": m["max_wave_speed"]}
    if "e2e" in m:
        line["e2e"] = m["e2e"]
    if "parity" in m:
        line["parity"] = m["parity"]
    del m
    torch.cuda.empty_cache()
    if not args.no_secondary and args.mode == "fused" and not args.rotate:
        # the other precisions / workloads of the path, short runs, so that they are measured by the same command
        from bench_subgrid import measure_subgrid
        sec = {}
        st = max(10, min(args.steps, 30))
        other = "f32" if dtype_name == "f64" else "f64"
        sec["unstructured_" + other] = short(measure_unstructured(args.level, other, st, args.warmup, device, e2e=False,
                                                                  parity=not args.no_parity, sampler_index=local))
        torch.cuda.empty_cache()
        for dn in ("f64", "f32"):
            sec["subgrid_" + dn] = measure_subgrid(6, dn, st, args.warmup, device, parity=not args.no_parity)
            torch.cuda.empty_cache()
        from bench_amr import amr_secondary
        sec["amr_c3"] = amr_secondary(dtype_name, 0, 1, device)
        sec["subgrid_amr_c4"] = amr_secondary(dtype_name, 0, 1, device, subgrid=True)
        from bench_hybrid import run_hybrid
        h = run_hybrid(argparse.Namespace(dtype=dtype_name, level=12, steps=st, warmup=args.warmup), 0, 1, device,
                       emit=False)
        sec["hybrid_c5"] = {"workload": h["config"]["workload"], "ms_per_step": h["ms_per_step"], "value": h["value"],
                            "roofline_frac": h["roofline"]["frac"], "parity": h["parity"]}
        line["secondary"] = sec
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
    ap.add_argument("--workload", default="unstructured", choices=["unstructured", "subgrid", "amr", "hybrid"],
                    help="unstructured = BASELINE configs[1] (the headline); subgrid = configs[3], Subgrid<4,4,4>; "
                         "amr = configs[2], adapt + repartition every --every steps; hybrid = configs[4], mixed "
                         "hex + prism + tet mesh, strong scaling (--level = tiles per direction, 648 * level^3 elements)")
    ap.add_argument("--every", type=int, default=10, help="amr: steps between adapts")
    ap.add_argument("--cycles", type=int, default=3, help="amr: adapt cycles")
    ap.add_argument("--check", action="store_true", help="amr: compare with a one-rank run of the same forest sequence")
    ap.add_argument("--subgrid", action="store_true", help="amr: Subgrid<4,4,4> elements (BASELINE configs[3] with AMR)")
    ap.add_argument("--level", type=int, default=None,
                    help="uniform refinement level per GPU (default 8 = 16.8M hexes; subgrid: 6 = 16.8M cells)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-timing comparison with oracle/_ref")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short fp32 / subgrid runs of the default line")
    ap.add_argument("--rotate", action="store_true", help="rotate the mesh in space: general-normal path of the plan")
    ap.add_argument("--mode", default="fused", choices=["fused", "unfused"],
                    help="unfused = the reference's schedule through the reference-shaped drop-in kernels")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.level is None and args.workload not in ("amr", "hybrid"):
        args.level = 6 if args.workload == "subgrid" else 8
    if args.impl == "reference":
        return run_reference(args)
    import t8gpu_b200
    t8gpu_b200.lib()  # fail loudly if the CUDA extension is missing
    run_ours(args)


if __name__ == "__main__":
    main()
