"""Throughput of the device remaps (SURVEY f-1) against the HBM roofline: adapt (refine everything / copy / coarsen
everything) for MeshManager elements and Subgrid<4,4,4> cells, and the partition remap (one rank, shifted mapping).
Bytes = variables read + written + adapt data; one JSON line per case."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import t8gpu_b200 as tb  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6551.0


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    best = 1e9
    for _ in range(reps):
        flush.zero_()                                  # L2 flush between repetitions
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def case(tag, dtype, n_old, n_new, ad, S, pk):
    dev = torch.device("cuda", 0)
    esz = 8 if dtype == torch.float64 else 4
    uo = [torch.rand(n_old * S, dtype=dtype, device=dev) for _ in range(5)]
    un = [torch.empty(n_new * S, dtype=dtype, device=dev) for _ in range(5)]
    vo, vn = torch.rand(n_old, dtype=dtype, device=dev), torch.empty(n_new, dtype=dtype, device=dev)
    ms = timed(lambda: tb.adapt_remap(ad, uo, un, vo, vn, subgrid_dim=3 if S == 64 else 0))
    bytes_ = 5 * esz * S * (min(n_old, n_new * 8 if n_new < n_old else n_old) + n_new) + esz * (n_old + n_new) + 4 * n_new
    print(json.dumps(dict(kernel="adapt_remap", case=tag, dtype=str(dtype).split(".")[-1], cells_old=n_old * S,
                          cells_new=n_new * S, ms=round(ms, 4), gbs=round(bytes_ / ms / 1e6, 1),
                          frac_of_hbm=round(bytes_ / ms / 1e6 / pk, 3))), flush=True)


def main():
    dev = torch.device("cuda", 0)
    pk = peak()
    for dtype in (torch.float64, torch.float32):
        for S, n in ((1, 1 << 21), (64, 1 << 15)):
            ar = torch.arange(8 * n + 1, dtype=torch.int32, device=dev)
            refine = torch.div(ar, 8, rounding_mode="floor").to(torch.int32)
            refine[-1] = n
            case("refine all (x8)", dtype, n, 8 * n, refine, S, pk)
            ident = torch.arange(8 * n + 1, dtype=torch.int32, device=dev)
            case("copy", dtype, 8 * n, 8 * n, ident, S, pk)
            coarsen = (torch.arange(n + 1, dtype=torch.int32, device=dev) * 8).to(torch.int32)
            case("coarsen all (/8)", dtype, 8 * n, n, coarsen, S, pk)
        # partition remap, one rank: new element e <- old element (e + shift) mod n
        n = 1 << 24
        esz = 8 if dtype == torch.float64 else 4
        old = [torch.rand(n, dtype=dtype, device=dev) for _ in range(5)]
        new = [torch.empty(n, dtype=dtype, device=dev) for _ in range(5)]
        vo, vn = torch.rand(n, dtype=dtype, device=dev), torch.empty(n, dtype=dtype, device=dev)
        ranks = torch.zeros(n, dtype=torch.int32, device=dev)
        idx = ((torch.arange(n, device=dev) + 12345) % n).to(torch.int32)
        tables = tb.RankTables([old], dev)
        vtab = torch.tensor([vo.data_ptr()], dtype=torch.int64).to(dev)
        ms = timed(lambda: tb.partition_remap(ranks, idx, new, tables, vn, vtab))
        bytes_ = 2 * 6 * esz * n + 8 * n
        print(json.dumps(dict(kernel="partition_remap", dtype=str(dtype).split(".")[-1], elements=n, ms=round(ms, 4),
                              gbs=round(bytes_ / ms / 1e6, 1), frac_of_hbm=round(bytes_ / ms / 1e6 / pk, 3))), flush=True)


if __name__ == "__main__":
    main()
