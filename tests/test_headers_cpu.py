"""The header-only mirror of the reference's template API (include/t8gpu/) compiles in a user-style translation unit,
for float and double, against the t8code / sc / MPI declarations of the t8mini shim (t8code is not installed here)."""
import os
import shutil
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_headers"))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_user_translation_unit_compiles():
    import build as hb
    hb.compile_check()


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
@pytest.mark.parametrize("src", ["mesh_harness.cu", "subgrid_harness.cu"])
def test_mesh_manager_translation_unit_compiles(src):
    inc = ["-I", os.path.join(HERE, "..", "include"), "-I", os.path.join(HERE, "..", "oracle", "ref_shim")]
    subprocess.check_call(["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                           "--expt-relaxed-constexpr", "-w"] + inc +
                          ["-c", os.path.join(HERE, "_headers", src), "-o", os.devnull])


@pytest.mark.skipif(shutil.which("nvcc") is None or not os.path.isdir("/root/reference"),
                    reason="needs nvcc and the reference sources")
def test_every_reference_example_translation_unit_compiles_against_the_mirror():
    """SURVEY 8(b) source compatibility: all nine translation units of the reference's examples (kernels, solvers and
    mains of examples/compressible_euler and examples/subgrid), unmodified, with -I include instead of the reference's
    own t8gpu/ headers, and the two solvers link into oracle/_ref/libmirror_*.so (exercised on the GPU by
    tests/test_mirror_gpu.py).  The build is cached by __graft_entry__.build()."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle import ref_build
    assert ref_build.build_mirror()
    objdir = os.path.join(ref_build.OUT, "obj_mirror")
    names = set(os.listdir(objdir))
    for tu in ref_build.EXAMPLE_TUS:
        base = os.path.basename(tu).replace(".cu", "")
        assert any(n.endswith(base + ".o") for n in names), tu
    assert ref_build.mirror_available()
