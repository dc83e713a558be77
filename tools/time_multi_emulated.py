#!/usr/bin/env python
"""Timing experiment: P = 2 ranks of the brick (2,1,1) emulated on ONE device (ghost reads go through the [var][rank]
tables into the other rank's buffer in LOCAL memory).  Separates the cost of the multi-rank code path of the stage
kernels from the cost of reading ghosts over NVLink: compare with tools/time_stage.py (one rank, no ghosts) and with
the 2-GPU bench."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import t8gpu_b200 as tb  # noqa: E402
from t8gpu_b200.solver import NB_STEPS, NVAR  # noqa: E402

level = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dtype = torch.float64 if (len(sys.argv) < 3 or sys.argv[2] == "f64") else torch.float32
dev = torch.device("cuda", 0)
P, brick = 2, (2, 1, 1)
conns = [tb.cartesian_uniform_connectivity(3, level, dtype, P, r, device=dev, brick=brick) for r in range(P)]
ns = [int(c["n_local"]) for c in conns]
bufs = []
for r in range(P):
    b = torch.zeros((NVAR * NB_STEPS + 1, ns[r]), dtype=dtype, device=dev)
    b[NVAR * NB_STEPS] = conns[r]["volumes"]
    tb.init_kelvin_helmholtz(3, conns[r]["centroids"], [b[k] for k in range(5)])
    bufs.append(b)
tabs = {s: tb.RankTables([[bufs[r][s * NVAR + k] for k in range(NVAR)] for r in range(P)], dev) for s in range(NB_STEPS)}
plans = [tb.Plan(tb.conn_to_host(c), dtype) for c in conns]
del conns
dt = 0.1 * 2.0 ** -level
nxt, prv = 0, 3


def step():
    global nxt, prv
    nxt, prv = prv, nxt
    for stage, sin, sout in ((1, prv, 1), (2, 1, 2), (3, 2, nxt)):
        for r in range(P):
            v = lambda s: [bufs[r][s * NVAR + k] for k in range(NVAR)]  # noqa: E731
            plans[r].stage(stage, v(sin), v(prv), v(sout), bufs[r][NVAR * NB_STEPS], dt, in_all=tabs[sin])


for _ in range(3):
    step()
torch.cuda.synchronize()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 10 / P)
print("emulated 2 ranks on one device, ms/step per rank:", ["%.3f" % r for r in res], plans[0].info)
