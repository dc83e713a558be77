"""Size-independent properties of the hot path, checked at BASELINE's full size (uniform periodic hex forest, level 8 =
16 777 216 elements, fp64) where the CPU oracle would take minutes, plus API edge cases (empty inputs, bad arguments)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _kh_solver(tb, level, dtype, device, mode):
    conn = tb.cartesian_uniform_connectivity(3, level, dtype, 1, 0, device=device)
    host = tb.conn_to_host(conn)
    sol = tb.EulerSolver(host, host["volumes"], dtype, device=device, mode=mode)
    tb.init_kelvin_helmholtz(3, conn["centroids"], sol.variables(sol.next))
    return sol


def test_full_size_conservation_determinism_and_two_paths(cuda):
    import t8gpu_b200 as tb
    level, dt = 8, 0.1 * 2.0 ** -8
    sol = _kh_solver(tb, level, torch.float64, cuda, "fused")
    n = sol.n
    assert n == 16777216 and sol.plan.info["n_chunks"] == 65536
    u0 = sol.state().clone()
    tot0 = u0.sum(1)          # uniform volumes: sum(vol * u) = vol * sum(u)
    for _ in range(3):
        sol.iterate(dt)
    u3 = sol.state().clone()
    assert torch.isfinite(u3).all()
    # conservation on the periodic mesh: every face flux enters two elements with opposite signs.  The truncated RK
    # literals of the reference (0.33333333333333 + 0.66666666666666 = 1 - 1e-14) take 1e-14 per step off every value
    tot3 = u3.sum(1)
    scale = u0.abs().sum(1).max()
    assert float(((tot3 - tot0).abs() / scale).max()) < 1e-13
    vmax = float(sol.max_wave_speed().item())
    assert 0.5 < vmax < 10.0
    # determinism: the same three steps from the same state give the same bits
    sol.set_state(u0)
    for _ in range(3):
        sol.iterate(dt)
    assert torch.equal(sol.state(), u3)
    assert float(sol.max_wave_speed().item()) == vmax
    del sol
    # the reference-shaped path (atomics, flux array in HBM) agrees with the fused one at full size
    ref = _kh_solver(tb, level, torch.float64, cuda, "unfused")
    for _ in range(3):
        ref.iterate(dt)
    err = float(((ref.state() - u3).abs().amax(1) / u3.abs().amax(1).clamp_min(1e-3 * float(u3.abs().max()))).max())
    assert err < 3e-12
    assert abs(float(ref.max_wave_speed().item()) - vmax) < 1e-12 * vmax


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_free_stream_is_preserved(cuda, dtype):
    """A constant state on a mesh with hanging faces stays constant: every element's face areas close (to rounding)
    and the two states of every face are identical, so only the reference's RK-literal decay remains."""
    import oracle
    import t8gpu_b200 as tb
    f = oracle.Forest(3, 3)
    lv, cent, vol, _ = f.elements()
    f = f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < 0.2, 20.0, 0.0), 10.0, 1, 4)
    lv, cent, vol, _ = f.elements()
    npdt = np.float64 if dtype == torch.float64 else np.float32
    conn = f.connectivity(dtype=npdt)
    u = np.tile(np.array([[1.3], [0.4], [-0.2], [0.1], [3.0]]), (1, f.num_elements)).astype(npdt)
    sol = tb.EulerSolver(conn, vol.astype(npdt), dtype, device=cuda, mode="fused")
    sol.set_state(u)
    for _ in range(10):
        sol.iterate(0.01)
    got = sol.state().cpu().numpy().astype(np.float64)
    tol = 1e-12 if dtype == torch.float64 else 2e-5
    assert np.abs(got / u.astype(np.float64) - 1).max() < tol


def test_walls_let_no_mass_through(cuda):
    """Closed box: total mass is conserved to rounding (the wall flux has no mass component), momentum is not."""
    import oracle
    import t8gpu_b200 as tb
    from util import perturbed_kh
    f = oracle.Forest(3, 3, periodic=False)
    conn = f.connectivity(dtype=np.float64)
    u0, vol = perturbed_kh(f, np.float64, seed=9)
    sol = tb.EulerSolver(conn, vol, torch.float64, device=cuda, mode="fused")
    sol.set_state(u0)
    for _ in range(5):
        sol.iterate(0.002)
    got = sol.state().cpu().numpy()
    m0, m1 = (u0[0] * vol).sum(), (got[0] * vol).sum()
    assert abs(m1 - m0) < 1e-13 * abs(m0)


def test_empty_rank_and_bad_arguments(cuda):
    import t8gpu_b200 as tb
    L = tb.lib()
    z64 = C.c_int64(0)
    # a rank without elements: plans build, stages / criteria / remaps are no-ops
    h = C.c_void_p()
    assert L.t8b200_plan_create(C.byref(h), 1, z64, z64, 0, 0, None, None, None, None, None, 0, None, None, None) == 0
    ptr5 = (C.c_void_p * 5)()
    assert L.t8b200_fused_stage_f64(h, 1, ptr5, None, None, ptr5, C.c_void_p(8), C.c_double(0.1), None, None) == 0
    assert L.t8b200_gradient_criteria_f64(h, C.c_void_p(8), None, C.c_void_p(8), C.c_void_p(8), None) == 0
    # wrong precision for the plan, bad stage, missing arrays
    assert L.t8b200_fused_stage_f32(h, 1, ptr5, None, None, ptr5, C.c_void_p(8), C.c_float(0.1), None, None) != 0
    assert L.t8b200_fused_stage_f64(h, 4, ptr5, None, None, ptr5, C.c_void_p(8), C.c_double(0.1), None, None) != 0
    assert L.t8b200_fused_stage_f64(h, 2, ptr5, None, None, ptr5, C.c_void_p(8), C.c_double(0.1), None, None) != 0
    assert L.t8b200_fused_stage_f64(None, 1, ptr5, None, None, ptr5, C.c_void_p(8), C.c_double(0.1), None, None) != 0
    L.t8b200_plan_destroy(h)
    L.t8b200_plan_destroy(None)
    assert L.t8b200_plan_create(C.byref(h), 1, C.c_int64(4), z64, 3, 0, None, None, None, None, None, 0, None, None,
                                None) != 0                                    # faces announced, arrays missing
    assert L.t8b200_plan_create(None, 1, z64, z64, 0, 0, None, None, None, None, None, 0, None, None, None) != 0
    sh = C.c_void_p()
    assert L.t8b200_subgrid_plan_create(C.byref(sh), 0, 3, z64, z64, 0, 0, None, None, None, None, None, None, None,
                                        None, 0, None, None, None, None, None) == 0
    assert L.t8b200_subgrid_fused_stage_f32(sh, 1, ptr5, None, None, ptr5, C.c_void_p(8), C.c_float(0.1), None) == 0
    L.t8b200_subgrid_plan_destroy(sh)
    assert L.t8b200_subgrid_plan_create(C.byref(sh), 0, 4, z64, z64, 0, 0, None, None, None, None, None, None, None,
                                        None, 0, None, None, None, None, None) != 0   # dim must be 2 or 3
    # a face that touches no local element is an inconsistent connectivity
    nbr = np.array([5, 6], np.int32)
    nrm = np.array([1.0, 0.0, 0.0])
    ar = np.array([1.0])
    rk = np.zeros(8, np.int32)
    ix = np.arange(8, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.t8b200_plan_create(C.byref(h), 1, C.c_int64(4), C.c_int64(4), 1, 0, p(nbr), p(nrm), p(ar), p(rk), p(ix), 0,
                                None, None, None) != 0
    # reference-shaped entry points
    assert L.t8b200_rk3_stage_f64(0, z64, 5, None, None, None, None, None, 1, C.c_double(0.1), None) != 0
    assert L.t8b200_max_speed_f64(None, z64, None, None) != 0
    out = torch.ones(1, dtype=torch.float64, device=cuda)
    assert L.t8b200_max_speed_f64(None, z64, C.c_void_p(out.data_ptr()), None) == 0
    torch.cuda.synchronize()
    assert float(out[0]) == 0.0


@pytest.mark.parametrize("scale", [1e-13, 1e13])
@pytest.mark.parametrize("mode", ["fused", "unfused"])
def test_fp32_dynamic_range(cuda, mode, scale):
    """Densities / pressures around 1e-13 and 1e+13 in fp32: the per-cell staging must not leave the float
    range where the reference's separate divides stay finite.  The equations are invariant under (rho, p) -> s (rho, p),
    so the step from the scaled state is compared with the fp32 oracle on the same scaled state."""
    import oracle
    import t8gpu_b200
    from util import TOL, perturbed_kh, rel_linf
    f = oracle.Forest(3, 3)
    conn = f.connectivity(dtype=np.float32)
    u0, vol = perturbed_kh(f, np.float32, seed=11)
    us = (u0.astype(np.float64) * scale).astype(np.float32)
    dt = 0.1 * 2.0 ** -3
    ref, _, _ = oracle.iterate(conn, vol, us, dt)
    assert np.isfinite(ref).all()
    sol = t8gpu_b200.EulerSolver(conn, vol, torch.float32, device=cuda, mode=mode)
    sol.set_state(us)
    sol.iterate(dt)
    got = sol.state().cpu().numpy()
    assert np.isfinite(got).all()
    assert rel_linf(got, ref) <= TOL[np.dtype(np.float32)]
