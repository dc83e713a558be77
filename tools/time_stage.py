#!/usr/bin/env python
"""Times K RK3 steps of the fused path on the level-L uniform hex forest without any checks (for ablation variants
selected with T8GPU_B200_LIB; see tools/build_variant.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import t8gpu_b200 as tb  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dtype = torch.float64 if (len(sys.argv) < 3 or sys.argv[2] == "f64") else torch.float32
dev = torch.device("cuda", 0)
conn = tb.cartesian_uniform_connectivity(3, level, dtype, 1, 0, device=dev)
host = tb.conn_to_host(conn)
sol = tb.EulerSolver(host, host["volumes"], dtype, device=dev, mode="fused")
tb.init_kelvin_helmholtz(3, conn["centroids"], sol.variables(sol.next))
u0 = sol.state().clone()
dt = 0.1 * 2.0 ** -level
res = []
for rep in range(3):
    sol.set_state(u0)
    for _ in range(3):
        sol.iterate(dt)
    sol.set_state(u0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        sol.iterate(dt)
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 10)
print(os.environ.get("T8GPU_B200_LIB", "default"), "ms/step", ["%.3f" % r for r in res])
L = tb.lib()
if hasattr(L, "t8b200_debug_phase_clocks"):
    import ctypes as C
    out = (C.c_ulonglong * 16)()
    L.t8b200_debug_phase_clocks(out, 1)
    sol.iterate(dt)
    L.t8b200_debug_phase_clocks(out, 0)
    names = ["phase0", "barrier1", "phase1", "barrier2", "phase2"]
    for o, w in ((0, "warp0"), (8, "warp7")):
        n = max(1, out[o + 5])
        print("  %s cycles/CTA:" % w, ", ".join("%s %.0f" % (nm, out[o + i] / n) for i, nm in enumerate(names)),
              "| total %.0f | face loops alone %.0f" % (sum(out[o:o + 5]) / n, out[o + 6] / n))
if hasattr(L, "t8b200_debug_cta_log"):
    import numpy as np
    log = np.zeros(4 * 65536, np.int64)
    L.t8b200_debug_cta_log(log.ctypes.data_as(C.c_void_p))
    log = log.reshape(-1, 4)
    for sm in (0, 57, 147):
        rows = log[log[:, 0] == sm]
        rows = rows[np.argsort(rows[:, 1])]
        t0 = rows[0, 1]
        print("  SM %d: %d CTAs; start / phase-1 start / end (cycles since first start):" % (sm, len(rows)))
        print("   ", " ".join("%d/%d/%d" % (r[1] - t0, r[2] - t0, r[3] - t0) for r in rows[:9]))
        mid = len(rows) // 2
        print("    mid-launch:", " ".join("%d/%d/%d" % (r[1] - t0, r[2] - t0, r[3] - t0) for r in rows[mid:mid + 9]))
