"""GPU parity of the refinement indicators: product (C ABI) vs the CPU oracle and vs the reference's own kernels."""
import numpy as np
import pytest
import torch

import oracle
from oracle import ref_cuda
from util import perturbed_kh

pytestmark = pytest.mark.gpu
DT = {np.float64: torch.float64, np.float32: torch.float32}
RTOL = {np.float64: 1e-13, np.float32: 1e-5}


def _forest(kind):
    if kind == "hex3":
        return oracle.Forest(3, 3)
    if kind == "hex3_walls":
        return oracle.Forest(3, 3, periodic=False)
    if kind == "quad5":
        return oracle.Forest(2, 5)
    f = oracle.Forest(3, 3)
    lv, cent, vol, _ = f.elements()
    return f.adapt(np.where(np.abs(cent[:, 2] - 0.5) < 0.2, 20.0, 0.0), 10.0, 1, 4)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind", ["hex3", "hex3_walls", "quad5", "hex_amr"])
def test_gradient_criteria_matches_oracle(cuda, kind, dtype):
    import t8gpu_b200 as tb
    f = _forest(kind)
    conn = f.connectivity(dtype=dtype)
    u0, vol = perturbed_kh(f, dtype, seed=3)
    ref = oracle.gradient_criteria(conn, u0[0], vol)
    plan = tb.Plan(conn, DT[dtype])
    got = tb.gradient_criteria(plan, torch.as_tensor(u0[0]).to(cuda), torch.as_tensor(vol).to(cuda)).cpu().numpy()
    assert np.abs(got - ref).max() <= RTOL[dtype] * np.abs(ref).max()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gradient_criteria_multi_rank(cuda, dtype):
    """3 ranks on one device: partition-boundary faces are summed by both owners from the other rank's density."""
    import t8gpu_b200 as tb
    f = _forest("hex_amr")
    u0, vol = perturbed_kh(f, dtype, seed=4)
    ref = oracle.gradient_criteria(f.connectivity(dtype=dtype), u0[0], vol)
    P = 3
    off = f.partition_offsets(P)
    rho = [torch.as_tensor(np.ascontiguousarray(u0[0, off[r]:off[r + 1]])).to(cuda) for r in range(P)]
    tab = torch.tensor([t.data_ptr() for t in rho], dtype=torch.int64, device=cuda)
    got = []
    for r in range(P):
        plan = tb.Plan(f.connectivity(P, r, dtype=dtype), DT[dtype])
        v = torch.as_tensor(np.ascontiguousarray(vol[off[r]:off[r + 1]])).to(cuda)
        got.append(tb.gradient_criteria(plan, rho[r], v, rho_all=tab).cpu().numpy())
    got = np.concatenate(got)
    assert np.abs(got - ref).max() <= RTOL[dtype] * np.abs(ref).max()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,level", [(3, 2), (2, 3)])
def test_subgrid_criteria_matches_oracle(cuda, dim, level, dtype):
    import t8gpu_b200 as tb
    f = oracle.Forest(dim, level)
    lv, cent, vol, _ = f.elements()
    S = 64 if dim == 3 else 16
    n = f.num_elements + 0
    rng = np.random.default_rng(8)
    rho = rng.uniform(1.0, 2.0, n * S).astype(dtype)
    ref = oracle.subgrid_criteria(dim, rho, vol.astype(dtype))
    got = tb.subgrid_criteria(dim, torch.as_tensor(rho).to(cuda), torch.as_tensor(vol.astype(dtype)).to(cuda)).cpu().numpy()
    assert np.abs(got - ref).max() <= 4 * RTOL[dtype] * np.abs(ref).max()
    # ragged: a number of elements that is not a multiple of the 32 per CTA, and none at all
    m = 37 if n > 37 else n - 1
    got = tb.subgrid_criteria(dim, torch.as_tensor(rho[:m * S]).to(cuda), torch.as_tensor(vol[:m].astype(dtype)).to(cuda))
    assert np.abs(got.cpu().numpy() - ref[:m]).max() <= 4 * RTOL[dtype] * np.abs(ref).max()
    assert tb.lib().t8b200_subgrid_criteria_f32(dim, 0, None, None, None, None) == 0


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,level", [(3, 3), (2, 4)])
def test_subgrid_criteria_bit_exact_vs_reference(cuda, dim, level, dtype):
    """Same loop order, same cbrt / sqrt, same contraction: the criteria (and hence the adapt decisions) are
    bit-identical to the reference kernel's."""
    import t8gpu_b200 as tb
    s = ref_cuda.RefSolver("sg", dtype, dim, level, True)
    conn = s.connectivity()
    u = s.get_state()          # the reference's own Kelvin-Helmholtz initial state
    s.iterate(0.1 * 2.0 ** -(level + 3), 3)
    u = s.get_state()
    ref = s.criteria()
    got = tb.subgrid_criteria(dim, torch.as_tensor(u[0]).to(cuda), torch.as_tensor(conn["volumes"]).to(cuda))
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gradient_criteria_vs_reference(cuda, dtype):
    """Reference: atomics in face order chosen by the hardware; here a fixed order: agreement to rounding."""
    import t8gpu_b200 as tb
    s = ref_cuda.RefSolver("uns", dtype, 3, 3, True)
    f = oracle.Forest(3, 3, True)
    u0, vol = perturbed_kh(f, dtype, seed=6)
    s.set_state(u0)
    ref = s.criteria()
    conn = s.connectivity()
    plan = tb.Plan(conn, DT[dtype])
    got = tb.gradient_criteria(plan, torch.as_tensor(u0[0]).to(cuda), torch.as_tensor(conn["volumes"]).to(cuda))
    assert np.abs(got.cpu().numpy() - ref).max() <= RTOL[dtype] * np.abs(ref).max()
