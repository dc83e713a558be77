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
