"""Generates the golden fixtures in this directory from the reference's OWN CUDA implementation (oracle/_ref, built by
oracle/ref_build.py from /root/reference).  Run on a GPU box:

    gpurun -- 'python tests/golden/make_golden.py && cp tests/golden/*.npz gpurun_out/'

The reference ships no tests or golden vectors; these are outputs of its kernels and of its mesh
managers (running over the t8mini forest facade) on small deterministic cases.  The -m "not gpu" suite pins the CPU
oracle against them; the -m gpu suite also compares the product with them.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from oracle import ref_cuda  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def conn_pack(c, subgrid=False):
    keys = ["ranks", "indices", "face_neighbors", "face_normals", "face_areas", "volumes"]
    if subgrid:
        keys += ["level_diff", "offsets"]
    out = {"conn_" + k: c[k] for k in keys}
    out["conn_counts"] = np.array([c["n_local"], c["n_ghost"], c["n_faces"], c["n_bfaces"]], np.int64)
    return out


def unstructured(name, dtype, dim, level, periodic, adapt, snaps, perturb):
    from util import perturbed_kh
    s = ref_cuda.RefSolver("uns", dtype, dim, level, periodic)
    f = oracle.Forest(dim, level, periodic)
    pack = {}
    if adapt:
        lv, cent, vol, _ = f.elements()
        crit = (np.where(np.abs(cent[:, dim - 1] - 0.5) < 0.2, 20.0, 0.0) +
                np.random.default_rng(5).uniform(0, 1, len(lv))).astype(dtype)
        s.mesh_adapt(crit)
        f = f.adapt(crit, 10.0, 1, 4)
        pack["adapt_crit"] = crit
    lv, cent, vol, _ = f.elements()
    if perturb:
        u0, _ = perturbed_kh(f, dtype, seed=31)
    else:
        u0 = oracle.init_kh_points(dim, cent.astype(dtype), dtype)
    s.set_state(u0)
    dt = 0.1 * 2.0 ** -(level + (1 if adapt else 0))
    pack.update(conn_pack(s.connectivity()))
    pack.update(u0=u0, dt=np.float64(dt), dim=dim, level=level, periodic=int(periodic), snaps=np.array(snaps))
    done = 0
    for k in snaps:
        s.iterate(dt, k - done)
        done = k
        pack["u_%d" % k] = s.get_state()
    pack["dt_cfl"] = np.float64(s.compute_timestep())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **pack)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in pack.items() if k.startswith("u_")},
          sha(pack["conn_face_neighbors"])[:12])


def subgrid(name, dtype, dim, level, adapt, snaps):
    s = ref_cuda.RefSolver("sg", dtype, dim, level, True)
    pack = {}
    if adapt:
        f = oracle.Forest(dim, level, True)
        lv, cent, vol, _ = f.elements()
        crit = (np.where(np.abs(cent[:, dim - 1] - 0.5) < 0.2, 1.0, 0.0)).astype(dtype)   # threshold 0.02
        pack["adapt_crit"] = crit
        pack["u_before_adapt"] = s.get_state()
        s.mesh_adapt(crit)
    u0 = s.get_state()               # the reference's own IC kernel (solver.inl:7-104), remapped if adapted
    dt = 0.1 * 2.0 ** -(level + (1 if adapt else 0) + 2)
    pack.update(conn_pack(s.connectivity(), subgrid=True))
    pack.update(u0=u0, dt=np.float64(dt), dim=dim, level=level, snaps=np.array(snaps))
    done = 0
    for k in snaps:
        s.iterate(dt, k - done)
        done = k
        pack["u_%d" % k] = s.get_state()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **pack)
    print(name, u0.shape, sha(pack["conn_face_neighbors"])[:12])


if __name__ == "__main__":
    assert ref_cuda.available(), "build oracle/_ref first (python oracle/ref_build.py)"
    for dt_, tag in ((np.float32, "f32"), (np.float64, "f64")):
        # BASELINE configs[0]: periodic quad level 6, KH, dt = 0.1 * 2^-6, 100 RK3 steps
        unstructured("uns_quad6_" + tag, dt_, 2, 6, True, False, [1, 10, 100], perturb=False)
        unstructured("uns_hex3_amr_walls_" + tag, dt_, 3, 3, False, True, [1, 5], perturb=True)
        subgrid("sg_hex2_" + tag, dt_, 3, 2, False, [1, 10])
        subgrid("sg_quad3_" + tag, dt_, 2, 3, False, [1, 10])
        subgrid("sg_hex2_amr_" + tag, dt_, 3, 2, True, [1, 5])
        subgrid("sg_quad3_amr_" + tag, dt_, 2, 3, True, [1, 5])
