"""Reference arm of bench.py (`--impl reference`).

The reference has NO CPU implementation of this path (flux + RK run only as CUDA kernels).  When oracle/_ref holds
the reference's own kernels compiled from /root/reference for sm_100a (oracle/ref_build.py), this arm times THOSE, with
the reference's own iterate() schedule, on the GPU of this box.  Otherwise it times the CPU oracle port on a bounded
sample.  Either way: same workload definition, metric and unit as the product arm."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))


def run_reference_arm(args):
    sys.path.insert(0, ROOT)
    try:
        from oracle import ref_cuda
        if ref_cuda.available():
            return ref_cuda.bench(args)
    except ImportError:
        pass
    t0 = time.time()
    subgrid = getattr(args, "workload", "unstructured") == "subgrid"
    if subgrid:
        from bench_subgrid import cpu_baseline
        n = 32768
    else:
        from bench import cpu_baseline
        n = 262144
    cb = cpu_baseline(args.dtype, budget_s=20.0)
    line = {"impl": "reference", "metric": "cell-updates/s per RK3 step", "value": cb["value"],
            "unit": "cell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * n / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": ("kelvin_helmholtz 3D Subgrid<4,4,4>, bounded sample: " if subgrid else
                                    "kelvin_helmholtz 3D uniform periodic hex mesh, bounded sample: ") + cb["sample"]},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line))
