"""Drop-in boundary, end to end: the reference's OWN example solvers (CompressibleEulerSolver,
SubgridCompressibleEulerSolver<Subgrid<4,4,4>> / <Subgrid<4,4>>; examples/*/{solver,kernels}*.cu compiled UNMODIFIED
from the reference sources) built twice --

    oracle/_ref/libref_*.so      against the reference's own t8gpu/ headers          (the reference)
    oracle/_ref/libmirror_*.so   against include/t8gpu/ of this repo + libt8gpu_b200 (the reference's solvers running on
                                 this repo's MemoryManager / MeshManager / SubgridMeshManager / SSP_3RK_step*)

-- and compared: connectivity arrays bit for bit, states over 20 steps, CFL time step, refinement criteria, the
adapt (+ partition) cycle, and what the save_*_to_vtk members hand to t8code.  Zero source edits on the reference side.
"""
import numpy as np
import pytest

import oracle
from oracle import ref_cuda
from util import TOL, perturbed_kh, rel_linf

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ref_cuda.available() and ref_cuda.mirror_available()),
                                 reason="oracle/_ref libref_* / libmirror_* not built (needs /root/reference)")]

CONN_KEYS = ("ranks", "indices", "face_neighbors", "face_normals", "face_areas", "volumes")


def pair(kind, dtype, dim, level, periodic):
    return (ref_cuda.RefSolver(kind, dtype, dim, level, periodic),
            ref_cuda.RefSolver(kind, dtype, dim, level, periodic, mirror=True))


def assert_same_mesh(r, m, subgrid):
    cr, cm = r.connectivity(), m.connectivity()
    for k in ("n_local", "n_ghost", "n_faces", "n_bfaces"):
        assert cr[k] == cm[k], k
    for k in CONN_KEYS + (("level_diff", "offsets") if subgrid else ()):
        assert cr[k].dtype == cm[k].dtype and np.array_equal(cr[k], cm[k]), k


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,level,periodic", [(3, 3, True), (3, 3, False), (2, 5, True)])
def test_reference_unstructured_solver_runs_on_the_mirror(cuda, dim, level, periodic, dtype):
    r, m = pair("uns", dtype, dim, level, periodic)
    assert_same_mesh(r, m, False)
    f = oracle.Forest(dim, level, periodic)
    u, _ = perturbed_kh(f, dtype, seed=31)
    r.set_state(u)
    m.set_state(u)
    dt = 0.1 * 2.0 ** -level
    for it in range(20):
        r.iterate(dt)
        m.iterate(dt)
        # same kernels, same arrays: only the hardware order of the flux atomics differs between two runs
        assert rel_linf(m.get_state(), r.get_state()) <= (it + 1) * TOL[np.dtype(dtype)], it
    assert abs(m.compute_timestep() - r.compute_timestep()) <= 1e-6 * r.compute_timestep()
    m.set_state(r.get_state())              # identical inputs from here on
    cr, cm = r.criteria(), m.criteria()     # estimate_gradient sums with atomics in hardware order
    assert np.abs(cm - cr).max() <= 10 * TOL[np.dtype(dtype)] * np.abs(cr).max()
    if dim == 3:   # MeshManager::adapt hard-codes the 3-D volume factors
        lv, cent, vol, _ = f.elements()
        crit = np.where(np.abs(cent[:, 2] - 0.5) < 0.2, 20.0, 0.0).astype(dtype)
        r.mesh_adapt(crit)
        m.mesh_adapt(crit)
        assert_same_mesh(r, m, False)
        assert np.array_equal(m.get_state(), r.get_state())   # remap: bit-exact
        for it in range(5):
            r.iterate(dt / 2)
            m.iterate(dt / 2)
        assert rel_linf(m.get_state(), r.get_state()) <= 5 * TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,level", [(3, 2), (3, 3), (2, 4)])
def test_reference_subgrid_solver_runs_on_the_mirror(cuda, dim, level, dtype):
    r, m = pair("sg", dtype, dim, level, True)
    assert_same_mesh(r, m, True)
    f = oracle.Forest(dim, level, True)
    lv, cent, vol, _ = f.elements()
    u = oracle.subgrid_init_kh(dim, cent.astype(dtype), lv, dtype)
    rng = np.random.default_rng(5)
    u = (u * (1 + 0.02 * rng.uniform(-1, 1, u.shape))).astype(dtype)
    r.set_state(u)
    m.set_state(u)
    dt = 0.1 * 2.0 ** -(level + 2)
    for it in range(20):
        r.iterate(dt)
        m.iterate(dt)
        assert rel_linf(m.get_state(), r.get_state()) <= (it + 1) * TOL[np.dtype(dtype)], it
    m.set_state(r.get_state())              # identical inputs from here on
    cr, cm = r.criteria(), m.criteria()
    assert np.array_equal(cr, cm)           # the reference's own criteria kernel on identical layouts
    # adapt + partition with the reference's rule (threshold 0.02), remap on the device, then keep stepping
    crit = np.where(np.abs(cent[:, 1] - 0.5) < 0.15, 1.0, 0.0).astype(dtype)
    r.mesh_adapt(crit)
    m.mesh_adapt(crit)
    assert_same_mesh(r, m, True)
    assert np.array_equal(m.get_state(), r.get_state())   # remap: bit-exact
    for it in range(5):
        r.iterate(dt / 2)
        m.iterate(dt / 2)
    assert rel_linf(m.get_state(), r.get_state()) <= 5 * TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_unstructured_vtk_fields(cuda, dtype):
    """save_conserved_variables_to_vtk (solver.cu:177-186): density, energy (scalars), momentum (interleaved vector),
    widened to double, handed to t8_forest_write_vtk_ext on the element forest."""
    r, m = pair("uns", dtype, 3, 3, True)
    f = oracle.Forest(3, 3, True)
    u, _ = perturbed_kh(f, dtype, seed=32)
    r.set_state(u)
    m.set_state(u)
    a, b = r.save("conserved", "ref_out"), m.save("conserved", "mirror_out")
    assert b["prefix"] == "mirror_out" and a["n_elements"] == b["n_elements"] == f.num_elements
    assert [n for n, _ in a["fields"]] == [n for n, _ in b["fields"]] == ["density", "energy", "momentum"]
    for (_, x), (_, y) in zip(a["fields"], b["fields"]):
        assert np.array_equal(x, y)
    assert np.array_equal(b["fields"][0][1], u[0].astype(np.float64))
    assert np.array_equal(b["fields"][2][1].reshape(-1, 3).T, u[1:4].astype(np.float64))


def morton_permutation(dim):
    """out[morton(i,j,k)] = in[i + 4 j + 16 k] (subgrid_mesh_manager.inl:1007-1049)."""
    S = 64 if dim == 3 else 16
    perm = np.zeros(S, np.int64)
    for flat in range(S):
        i, j, k = flat & 3, (flat >> 2) & 3, flat >> 4
        mo = 0
        for l in range(2):
            mo |= ((i >> l) & 1) << (dim * l) | ((j >> l) & 1) << (dim * l + 1)
            if dim == 3:
                mo |= ((k >> l) & 1) << (dim * l + 2)
        perm[mo] = flat
    return perm


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim,level", [(3, 2), (2, 3)])
def test_subgrid_vtk_cell_data_in_z_order(cuda, dim, level, dtype):
    """save_density_to_vtk -> SubgridMeshManager::save_variable_to_vtk (subgrid_mesh_manager.inl:1051-1124): the forest
    refined twice more (every cell a leaf) and the density permuted from column-major to that forest's Morton order;
    save_mesh_to_vtk: the element forest, no data."""
    r, m = pair("sg", dtype, dim, level, True)
    f = oracle.Forest(dim, level, True)
    lv, cent, vol, _ = f.elements()
    S = 64 if dim == 3 else 16
    rng = np.random.default_rng(6)
    u = oracle.subgrid_init_kh(dim, cent.astype(dtype), lv, dtype)
    u[0] = (u[0] * (1 + 0.1 * rng.uniform(-1, 1, u.shape[1]))).astype(dtype)   # every cell its own density
    r.set_state(u)
    m.set_state(u)
    a, b = r.save("density", "ref_rho"), m.save("density", "mirror_rho")
    assert a["n_elements"] == b["n_elements"] == f.num_elements * S
    assert b["min_level"] == b["max_level"] == level + 2 and b["prefix"] == "mirror_rho"
    assert len(b["fields"]) == 1 and b["fields"][0][0] == a["fields"][0][0] == "variables"
    assert np.array_equal(a["fields"][0][1], b["fields"][0][1])
    want = u[0].astype(np.float64).reshape(-1, S)[:, morton_permutation(dim)].reshape(-1)
    assert np.array_equal(b["fields"][0][1], want)
    a, b = r.save("mesh", "ref_mesh"), m.save("mesh", "mirror_mesh")
    assert a["n_elements"] == b["n_elements"] == f.num_elements and len(b["fields"]) == 0
    # the manager's own forest is untouched by the output refinement: stepping still works
    m.iterate(1e-4)
    assert np.isfinite(m.get_state()).all()
