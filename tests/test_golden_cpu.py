"""Pins the CPU oracle against golden outputs of the reference's OWN CUDA implementation (tests/golden/*.npz, generated
on a B200 by tests/golden/make_golden.py from oracle/_ref).  Runs without a GPU."""
import glob
import os

import numpy as np
import pytest

import oracle
from util import TOL, rel_linf

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def forest_for(g, subgrid_threshold=None):
    f = oracle.Forest(int(g["dim"]), int(g["level"]), bool(int(g["periodic"])) if "periodic" in g else True)
    if "adapt_crit" in g:
        if subgrid_threshold is None:
            f = f.adapt(g["adapt_crit"], 10.0, 1, 4)             # MeshManager: b = 10, levels 1..4
        else:
            f = f.adapt(g["adapt_crit"], subgrid_threshold, 1, 6)  # SubgridMeshManager: b = 0.02, levels 1..6
    return f


def check_conn(g, conn, keys):
    cnt = g["conn_counts"]
    assert (conn["n_local"], conn["n_ghost"], conn["n_faces"], conn["n_bfaces"]) == tuple(cnt)
    for k in keys:
        assert g["conn_" + k].dtype == conn[k].dtype, k
        assert np.array_equal(g["conn_" + k], conn[k]), k


def test_all_fixtures_present():
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "*.npz")))
    assert len(names) == 12, names


@pytest.mark.parametrize("tag,dtype", [("f32", np.float32), ("f64", np.float64)])
@pytest.mark.parametrize("case", ["uns_quad6", "uns_hex3_amr_walls"])
def test_unstructured_oracle_vs_reference_golden(case, tag, dtype):
    g = load(case + "_" + tag)
    f = forest_for(g)
    conn = f.connectivity(dtype=dtype)
    # connectivity + volumes: bit exact against the reference's MeshManager (over t8mini)
    check_conn(g, conn, ("ranks", "indices", "face_neighbors", "face_normals", "face_areas"))
    vol = f.elements()[2].astype(dtype)
    assert np.array_equal(g["conn_volumes"], vol)
    # arithmetic: advance the oracle from the golden initial state
    u, dt, done = g["u0"], float(g["dt"]), 0
    sp = np.zeros(conn["n_faces"] + conn["n_bfaces"], dtype)
    for k in g["snaps"]:
        for _ in range(int(k) - done):
            u, _, _ = oracle.iterate(conn, vol, u, dt, speed=sp)
        done = int(k)
        err = rel_linf(u, g["u_%d" % k])
        assert err <= done * TOL[np.dtype(dtype)], (case, tag, k, err)
    assert abs(oracle.compute_timestep(sp, dtype(0.7), 4) - float(g["dt_cfl"])) <= 1e-5 * float(g["dt_cfl"])


@pytest.mark.parametrize("tag,dtype", [("f32", np.float32), ("f64", np.float64)])
@pytest.mark.parametrize("case", ["sg_hex2", "sg_quad3", "sg_hex2_amr", "sg_quad3_amr"])
def test_subgrid_oracle_vs_reference_golden(case, tag, dtype):
    g = load(case + "_" + tag)
    f = forest_for(g, subgrid_threshold=0.02)
    dim = int(g["dim"])
    conn = f.connectivity(subgrid=True, dtype=dtype)
    check_conn(g, conn, ("ranks", "indices", "face_neighbors", "face_normals", "face_areas", "level_diff", "offsets"))
    lv, cent, vol, _ = f.elements()
    vol = vol.astype(dtype)
    assert np.array_equal(g["conn_volumes"], vol)
    if "adapt_crit" not in g:
        # the reference's own IC kernel (solver.inl:7-104) against the restated one
        u_ic = oracle.subgrid_init_kh(dim, cent.astype(dtype), lv, dtype)
        assert rel_linf(u_ic, g["u0"]) <= 8 * np.finfo(dtype).eps
    u, dt, done = g["u0"], float(g["dt"]), 0
    for k in g["snaps"]:
        for _ in range(int(k) - done):
            u, _, _ = oracle.subgrid_iterate(conn, vol, u, dt)
        done = int(k)
        err = rel_linf(u, g["u_%d" % k])
        assert err <= done * TOL[np.dtype(dtype)], (case, tag, k, err)
