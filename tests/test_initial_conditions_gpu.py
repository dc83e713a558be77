"""a25: the Kelvin-Helmholtz state on the globe of the unstructured example (examples/compressible_euler/solver.cu:17-72).
The oracle's restatement is pinned against the reference's own constructor (oracle/_ref: the reference solver runs its
host lambda over the t8code stand-in's centroids); the device kernel is compared with the oracle."""
import numpy as np
import pytest
import torch

import oracle
from oracle import ref_cuda

pytestmark = pytest.mark.gpu
DT = {np.float32: torch.float32, np.float64: torch.float64}


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oracle_restatement_equals_the_reference_constructor(cuda, dtype):
    s = ref_cuda.RefSolver("uns", dtype, 3, 3, True)
    ref = s.get_state()
    cent = oracle.Forest(3, 3).elements()[1]
    got = oracle.init_spherical_kh_points(cent.astype(dtype), dtype)
    assert ref.shape == got.shape
    # the same libm on the same host; a few ulp are left for the contraction choices of the two host compilers
    assert np.abs(ref.astype(np.float64) - got.astype(np.float64)).max() <= 8 * np.finfo(dtype).eps * np.abs(ref).max()
    assert np.array_equal(ref[0], got[0])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_device_kernel_matches_the_oracle(cuda, dtype):
    import t8gpu_b200 as tb
    rng = np.random.default_rng(3)
    n = 20000
    c = rng.uniform(-1.0, 1.0, (n, 3))
    c *= (0.6 + 0.4 * rng.random((n, 1))) / np.linalg.norm(c, axis=1, keepdims=True)      # a shell 0.6 <= r <= 1
    c = np.concatenate([c, oracle.Forest(3, 3).elements()[1]]).astype(dtype)
    want = oracle.init_spherical_kh_points(c, dtype)
    u = [torch.empty(c.shape[0], dtype=DT[dtype], device=cuda) for _ in range(5)]
    tb.init_spherical_kelvin_helmholtz(torch.as_tensor(c).to(cuda), u)
    got = torch.stack(u).cpu().numpy()
    assert np.isfinite(got).all() and np.array_equal(got[0], want[0])            # the density jump is exact
    tol = 1e-13 if dtype == np.float64 else 5e-6
    assert np.abs(got.astype(np.float64) - want.astype(np.float64)).max() <= tol * np.abs(want).max()
    # empty input, missing arrays
    L = tb.lib()
    import ctypes as C
    assert L.t8b200_init_spherical_kelvin_helmholtz_f64(C.c_int64(0), None, None, None) == 0
    assert L.t8b200_init_spherical_kelvin_helmholtz_f64(C.c_int64(4), None, None, None) != 0
