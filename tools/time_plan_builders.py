"""Build time of the tile plan by builder (SURVEY f-2): three-kernel device builder (structured-only meshes), generic
device builder (one CUDA thread per block of 256 elements, csrc/plan_block.cuh) and the multithreaded host builder
(includes the device -> host copy of the connectivity it needs).  Prints one JSON line per mesh."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import t8gpu_b200 as tb  # noqa: E402


def timed(fn, reps=3):
    best = 1e9
    out = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.time()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.time() - t)
        if _ + 1 < reps:
            out = None
    return best, out


def run(tag, conn, dtype):
    line = dict(mesh=tag, elements=int(conn["n_local"]), faces=int(conn["n_faces"]), dtype=str(dtype).split(".")[-1])
    os.environ.pop("T8B200_DEVICE_PLAN", None)
    t, p = timed(lambda: tb.Plan.from_device(conn, dtype))
    line["device_s"], line["chunks"] = round(t, 4), int(p.info["n_chunks"])
    del p
    os.environ["T8B200_DEVICE_PLAN"] = "generic"
    t, p = timed(lambda: tb.Plan.from_device(conn, dtype))
    line["device_generic_s"] = round(t, 4)
    del p
    os.environ.pop("T8B200_DEVICE_PLAN", None)
    t, p = timed(lambda: tb.Plan(tb.conn_to_host(conn), dtype), reps=2)
    line["host_s"], line["host_threads"] = round(t, 4), os.cpu_count()
    print(json.dumps(line), flush=True)


def main():
    dev = torch.device("cuda", 0)
    dtype = torch.float64
    run("uniform periodic hex level 8", tb.cartesian_uniform_connectivity(3, 8, dtype, 1, 0, device=dev), dtype)
    import oracle                                         # host forest (t8code stand-in) for the adaptive mesh only
    f = oracle.Forest(3, 5)
    for width, top in ((0.25, 6), (0.15, 7)):
        lv, cent, vol, _ = f.elements()
        f = f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < width, 20.0, 0.0), 10.0, 1, top)
    lv, cent, vol, _ = f.elements()
    conn = tb.forest_connectivity(3, True, tb.morton_keys(3, lv, cent), lv, dtype, device=dev)
    run("adaptive periodic hex levels %d..%d" % (lv.min(), lv.max()), conn, dtype)


if __name__ == "__main__":
    main()
