"""north_star: conserved variables within relative L-infinity 1e-12 (fp64) / 1e-5 (fp32) PER STEP over 100 steps.
Every case restarts the checker from the product's state before each step, so the bound asserted is the per-step one
(no accumulation allowance): unstructured fused / reference-shaped kernels, fp64 / fp32, quad (configs[0]) and hex
meshes, Subgrid<4,4,4> and Subgrid<4,4>; plus the product against the golden snapshots of the reference's own kernels
(steps 1 / 10 / 100) and against the live reference at a size close to the measured one."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import ref_cuda
from util import TOL, rel_linf

pytestmark = pytest.mark.gpu
DT = {np.float64: torch.float64, np.float32: torch.float32}
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("mode", ["fused", "unfused"])
@pytest.mark.parametrize("dim,level", [(2, 6), (3, 4)])
def test_hundred_steps_unstructured(cuda, dim, level, mode, dtype):
    """(2, 6): the level-6 quad mesh of configs[0]; (3, 4): 4 096 hexahedra = 16 structured chunks (the kernel of the
    headline configuration).  Unperturbed Kelvin-Helmholtz data, the reference's dt = 0.1 * 2^-level."""
    import t8gpu_b200
    forest = oracle.Forest(dim, level)
    conn = forest.connectivity(dtype=dtype)
    lv, cent, vol, _ = forest.elements()
    vol = vol.astype(dtype)
    sol = t8gpu_b200.EulerSolver(conn, vol, DT[dtype], device=cuda, mode=mode)
    sol.set_state(oracle.init_kh_points(dim, cent.astype(dtype), dtype))
    dt = 0.1 * 2.0 ** -level
    worst = 0.0
    for it in range(100):
        cur = sol.state().cpu().numpy().copy()
        ref, _, _ = oracle.iterate(conn, vol, cur, dt)
        sol.iterate(dt)
        worst = max(worst, rel_linf(sol.state().cpu().numpy(), ref))
    assert worst <= TOL[np.dtype(dtype)], worst


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("mode", ["fused", "unfused"])
@pytest.mark.parametrize("dim,level", [(3, 2), (2, 3)])
def test_hundred_steps_subgrid(cuda, dim, level, mode, dtype):
    """Subgrid<4,4,4> (level 2: 64 elements = 4 096 cells, structured chunks) and Subgrid<4,4>, the reference's own
    subgrid problem scaled down: IC of examples/subgrid/solver.inl, dt = 0.1 * 2^-(level + 2)."""
    import t8gpu_b200
    if mode == "fused" and dim == 2:
        pytest.skip("the fused subgrid stage is the 3-D one (SubgridEulerSolver)")
    forest = oracle.Forest(dim, level)
    conn = forest.connectivity(subgrid=True, dtype=dtype)
    lv, cent, vol, _ = forest.elements()
    vol = vol.astype(dtype)
    sol = t8gpu_b200.SubgridEulerSolver(conn, vol, DT[dtype], device=cuda, mode=mode)
    sol.set_state(oracle.subgrid_init_kh(dim, cent.astype(dtype), lv, dtype))
    dt = 0.1 * 2.0 ** -(level + 2)
    worst = 0.0
    for it in range(100):
        cur = sol.state().cpu().numpy().copy()
        ref, _, _ = oracle.subgrid_iterate(conn, vol, cur, dt)
        sol.iterate(dt)
        worst = max(worst, rel_linf(sol.state().cpu().numpy(), ref))
    assert worst <= TOL[np.dtype(dtype)], worst


@pytest.mark.parametrize("tag,dtype", [("f32", np.float32), ("f64", np.float64)])
@pytest.mark.parametrize("mode", ["fused", "unfused"])
@pytest.mark.parametrize("case", ["uns_quad6", "uns_hex3_amr_walls"])
def test_product_vs_reference_golden_snapshots(cuda, case, mode, tag, dtype):
    """The product advanced from the golden initial state against the states the reference's OWN kernels produced
    after 1 / 10 / 100 steps (tests/golden/*.npz; uns_quad6 = configs[0]), and the CFL time step."""
    import t8gpu_b200
    g = np.load(os.path.join(GOLD, "%s_%s.npz" % (case, tag)))
    f = oracle.Forest(int(g["dim"]), int(g["level"]), bool(int(g["periodic"])) if "periodic" in g else True)
    if "adapt_crit" in g:
        f = f.adapt(g["adapt_crit"], 10.0, 1, 4)
    conn = f.connectivity(dtype=dtype)
    vol = f.elements()[2].astype(dtype)
    sol = t8gpu_b200.EulerSolver(conn, vol, DT[dtype], device=cuda, mode=mode)
    sol.set_state(g["u0"])
    dt, done = float(g["dt"]), 0
    for k in g["snaps"]:
        for _ in range(int(k) - done):
            sol.iterate(dt)
        done = int(k)
        err = rel_linf(sol.state().cpu().numpy(), g["u_%d" % k])
        # two independent runs (the golden one sums its fluxes with atomics in hardware order): k steps of drift
        assert err <= done * TOL[np.dtype(dtype)], (case, tag, k, err)
    assert abs(sol.compute_timestep() - float(g["dt_cfl"])) <= 1e-5 * float(g["dt_cfl"])


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_level7_hex_against_the_live_reference_per_step(cuda, dtype):
    """2 097 152 hexahedra (one refinement level below the measured configuration, whose own comparison runs inside
    bench.py): per step from the reference's state, 5 steps, structured-chunk kernel, plan built on the device."""
    import t8gpu_b200 as tb
    level = 7
    conn = tb.cartesian_uniform_connectivity(3, level, DT[dtype], 1, 0, device=cuda)
    plan = tb.Plan.from_device(conn, DT[dtype])
    assert plan is not None
    sol = tb.EulerSolver(dict(n_local=int(conn["n_local"]), n_faces=int(conn["n_faces"]), n_bfaces=0), conn["volumes"],
                         DT[dtype], device=cuda, plan=plan)
    tb.init_kelvin_helmholtz(3, conn["centroids"], sol.variables(sol.next))
    del conn
    ref = ref_cuda.RefSolver("uns", dtype, 3, level, True)
    ref.set_state(sol.state().cpu().numpy())
    dt = 0.1 * 2.0 ** -level
    for it in range(5):
        sol.set_state(ref.get_state())          # per-step: both advance from the reference's state
        ref.iterate(dt)
        sol.iterate(dt)
        err = rel_linf(sol.state().cpu().numpy(), ref.get_state())
        assert err <= TOL[np.dtype(dtype)], (it, err)
    ref.close()
