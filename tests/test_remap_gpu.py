"""GPU parity of the device-side remap after adapt / partition: product (C ABI) vs the CPU oracle, bit for bit
(copies and injections are exact; the coarsening means are summed in the reference's order), and vs the reference's
own adapt() (oracle/_ref)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import ref_cuda

pytestmark = pytest.mark.gpu
DT = {np.float64: torch.float64, np.float32: torch.float32}


def _adapted(dim, level, seed, subgrid):
    """A forest, a criterion that refines some elements and coarsens others, the adapted forest and the index map."""
    f = oracle.Forest(dim, level)
    lv, cent, vol, _ = f.elements()
    rng = np.random.default_rng(seed)
    hi, lo = (1.0, 0.0) if subgrid else (20.0, 0.0)
    thr = 0.02 if subgrid else 10.0
    crit = np.where(np.abs(cent[:, dim - 1] - 0.5) < 0.2, hi, lo)
    f1 = f.adapt(crit, thr, 1, level + 1)                      # refine a slab
    lv1, cent1, vol1, _ = f1.elements()
    crit1 = np.where(np.abs(cent1[:, 0] - 0.5) < 0.25, hi, lo)  # then refine another one, coarsen the rest
    f2 = f1.adapt(crit1, thr, 1, level + 1)
    return f1, vol1, f2, f1.adapt_map(f2), rng


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,level", [(3, 2), (2, 3)])
def test_adapt_remap_elements(cuda, dim, level, dtype):
    import t8gpu_b200 as tb
    f1, vol1, f2, ad, rng = _adapted(dim, level, 3, False)
    assert (np.diff(ad) == 0).any() and (np.diff(ad) > 1).any() and (np.diff(ad) == 1).any()
    n_old, n_new = f1.num_elements, f2.num_elements
    u_old = rng.uniform(0.5, 2.0, (5, n_old)).astype(dtype)
    vol_old = vol1.astype(dtype)
    ref_u, ref_v = oracle.adapt_remap(ad, u_old, vol_old, 0)
    uo = torch.as_tensor(u_old).to(cuda)
    un = torch.full((5, n_new), -7.0, dtype=DT[dtype], device=cuda)
    vo = torch.as_tensor(vol_old).to(cuda)
    vn = torch.zeros(n_new, dtype=DT[dtype], device=cuda)
    tb.adapt_remap(torch.as_tensor(ad).to(cuda), list(uo), list(un), vo, vn, 0)
    assert np.array_equal(un.cpu().numpy(), ref_u)
    assert np.array_equal(vn.cpu().numpy(), ref_v)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,level", [(3, 2), (2, 3)])
def test_adapt_remap_subgrid(cuda, dim, level, dtype):
    import t8gpu_b200 as tb
    f1, vol1, f2, ad, rng = _adapted(dim, level, 5, True)
    assert (np.diff(ad) == 0).any() and (np.diff(ad) > 1).any() and (np.diff(ad) == 1).any()
    S = 64 if dim == 3 else 16
    n_old, n_new = f1.num_elements, f2.num_elements
    u_old = rng.uniform(0.5, 2.0, (5, n_old * S)).astype(dtype)
    vol_old = vol1.astype(dtype)
    ref_u, ref_v = oracle.adapt_remap(ad, u_old, vol_old, dim)
    uo = torch.as_tensor(u_old).to(cuda)
    un = torch.full((5, n_new * S), -7.0, dtype=DT[dtype], device=cuda)
    vo = torch.as_tensor(vol_old).to(cuda)
    vn = torch.zeros(n_new, dtype=DT[dtype], device=cuda)
    tb.adapt_remap(torch.as_tensor(ad).to(cuda), list(uo), list(un), vo, vn, dim)
    assert np.array_equal(un.cpu().numpy(), ref_u)
    assert np.array_equal(vn.cpu().numpy(), ref_v)


@pytest.mark.parametrize("cpe", [1, 64, 16])
def test_partition_remap(cuda, cpe):
    """Repartition of 3 ranks' data into a new SFC split: gather through [var][rank] tables."""
    import t8gpu_b200 as tb
    rng = np.random.default_rng(9)
    n_old = [37, 64, 20]
    u_old = [rng.uniform(-1, 1, (5, n * cpe)) for n in n_old]
    v_old = [rng.uniform(1, 2, n) for n in n_old]
    # the new rank receives a contiguous range of the global order that straddles all three old ranks
    glob = [(r, i) for r, n in enumerate(n_old) for i in range(n)]
    take = glob[30:110]
    ranks = np.array([r for r, _ in take], np.int32)
    indices = np.array([i for _, i in take], np.int32)
    ref_u, ref_v = oracle.partition_remap(ranks, indices, u_old, v_old, cpe)
    d_u = [torch.as_tensor(u).to(cuda) for u in u_old]
    d_v = [torch.as_tensor(v).to(cuda) for v in v_old]
    tabs = tb.RankTables([list(u) for u in d_u], cuda)
    vtab = torch.tensor([v.data_ptr() for v in d_v], dtype=torch.int64, device=cuda)
    un = torch.zeros((5, len(take) * cpe), dtype=torch.float64, device=cuda)
    vn = torch.zeros(len(take), dtype=torch.float64, device=cuda)
    tb.partition_remap(torch.as_tensor(ranks).to(cuda), torch.as_tensor(indices).to(cuda), list(un), tabs, vn, vtab, cpe)
    assert np.array_equal(un.cpu().numpy(), ref_u)
    assert np.array_equal(vn.cpu().numpy(), ref_v)


def test_remap_empty_and_bad_arguments(cuda):
    import ctypes as C
    import t8gpu_b200 as tb
    L = tb.lib()
    assert L.t8b200_adapt_remap_f64(0, 5, C.c_int64(0), None, None, None, None, None, None) == 0
    assert L.t8b200_adapt_remap_f64(1, 5, C.c_int64(4), None, None, None, None, None, None) != 0   # bad subgrid_dim
    assert L.t8b200_adapt_remap_f64(0, 9, C.c_int64(4), None, None, None, None, None, None) != 0   # too many variables
    assert L.t8b200_partition_remap_f32(5, C.c_int64(0), 1, None, None, None, None, None, None, None) == 0
    assert L.t8b200_partition_remap_f32(5, C.c_int64(3), 0, None, None, None, None, None, None, None) != 0


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("kind,dim,level", [("uns", 3, 3), ("sg", 3, 2), ("sg", 2, 3)])
def test_adapt_remap_vs_reference_adapt(cuda, kind, dim, level, dtype):
    """The reference's own MeshManager / SubgridMeshManager::adapt (running over t8mini) against the product remap
    driven by the oracle's index map."""
    import t8gpu_b200 as tb
    s = ref_cuda.RefSolver(kind, dtype, dim, level, True)
    f = oracle.Forest(dim, level, True)
    lv, cent, vol, _ = f.elements()
    S = s.cells_per_element
    rng = np.random.default_rng(2)
    u_old = rng.uniform(0.5, 2.0, (5, f.num_elements * S)).astype(dtype)
    s.set_state(u_old)
    hi, thr = (20.0, 10.0) if kind == "uns" else (1.0, 0.02)
    crit = np.where(np.abs(cent[:, dim - 1] - 0.5) < 0.2, hi, 0.0).astype(dtype)
    s.mesh_adapt(crit)
    f2 = f.adapt(crit, thr, 1, 4 if kind == "uns" else 6)
    ad = f.adapt_map(f2)
    assert s.counts()["n_local"] == f2.num_elements
    uo = torch.as_tensor(u_old).to(cuda)
    un = torch.zeros((5, f2.num_elements * S), dtype=DT[dtype], device=cuda)
    vo = torch.as_tensor(vol.astype(dtype)).to(cuda)
    vn = torch.zeros(f2.num_elements, dtype=DT[dtype], device=cuda)
    tb.adapt_remap(torch.as_tensor(ad).to(cuda), list(uo), list(un), vo, vn, 0 if kind == "uns" else dim)
    assert np.array_equal(un.cpu().numpy(), s.get_state())
    assert np.array_equal(vn.cpu().numpy(), s.connectivity()["volumes"])
