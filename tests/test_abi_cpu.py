"""The C-ABI library loads without a GPU and exports every symbol include/t8gpu_b200.h declares."""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "t8gpu_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(t8b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import t8gpu_b200
    from t8gpu_b200 import build
    build.build()
    lib = C.CDLL(t8gpu_b200.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 16
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.t8b200_version() >= 100


def test_argument_errors_do_not_need_a_gpu():
    import t8gpu_b200
    lib = t8gpu_b200.lib()
    # cudaErrorInvalidValue == 1; argument validation happens before any CUDA call
    assert lib.t8b200_rk3_stage_f64(7, C.c_int64(10), 5, None, None, None, None, None, 1, C.c_double(0.1), None) == 1
    assert lib.t8b200_plan_create(None, 1, C.c_int64(0), C.c_int64(0), 0, 0, None, None, None, None, None, 0, None,
                                  None, None) == 1
    assert lib.t8b200_plan_info(None, None) == 1


def test_no_product_code_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "t8gpu_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M) or "liboracle" in txt or \
                        "oracle/" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
