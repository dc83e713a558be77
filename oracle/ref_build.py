"""TEST INFRASTRUCTURE: builds oracle/_ref/ -- the reference's OWN hot-path translation units compiled for sm_100a
from where they lie under /root/reference (never copied into the repo), linked against the t8mini shim
(oracle/ref_shim) and the C harnesses (oracle/ref_harness).

    libref_uns_f32.so / libref_uns_f64.so   examples/compressible_euler/{kernels,solver}.cu
    libref_sg_f32.so  / libref_sg_f64.so    examples/subgrid/{kernels,solver}_{2d,3d}.cu

fp64: the reference hard-codes `float_type = float` (t8gpu/memory/memory_manager.h:29,39).  The f64 variants are
compiled against a scratch copy of that one header with `= float;` -> `= double;`, generated at build time into
oracle/_ref/inc_f64/ (git-ignored) and placed first on the include path; every other file is used in place.
The reference's own build system (CMake + find_package(T8CODE)) is not run.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "ref_shim")
HARN = os.path.join(HERE, "ref_harness")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-lineinfo", "--expt-relaxed-constexpr",
         "-rdc=true", "-Xcompiler", "-fPIC", "-w"]

LIBS = {
    "uns": (["examples/compressible_euler/kernels.cu", "examples/compressible_euler/solver.cu"], "harness_uns.cu",
            "examples/compressible_euler"),
    "sg": (["examples/subgrid/kernels_2d.cu", "examples/subgrid/kernels_3d.cu", "examples/subgrid/solver_2d.cu",
            "examples/subgrid/solver_3d.cu"], "harness_sg.cu", "examples/subgrid"),
}


# The same reference translation units compiled against the PRODUCT's header mirror (include/t8gpu/) instead of the
# reference's own t8gpu/ headers: the reference's unmodified example solvers running on this repo's managers.
#     libmirror_uns_f32.so / _f64.so, libmirror_sg_f32.so / _f64.so   (same harness, same C symbols as libref_*)
MIRROR_INC = os.path.join(os.path.dirname(HERE), "include")
PRODUCT_LIBDIR = os.path.join(os.path.dirname(HERE), "t8gpu_b200")
# every translation unit of the reference's examples (compile check of the mirror: SURVEY 8(b) source compatibility)
EXAMPLE_TUS = ["examples/compressible_euler/kernels.cu", "examples/compressible_euler/solver.cu",
               "examples/compressible_euler/main.cu", "examples/subgrid/kernels_2d.cu", "examples/subgrid/kernels_3d.cu",
               "examples/subgrid/solver_2d.cu", "examples/subgrid/solver_3d.cu", "examples/subgrid/main_2d.cu",
               "examples/subgrid/main_3d.cu"]


def mirror_available():
    return all(os.path.exists(os.path.join(OUT, "libmirror_%s_%s.so" % (k, p))) for k in LIBS for p in ("f32", "f64"))


def _run_jobs(jobs, verbose=False, width=6):
    running, pending = [], list(jobs)
    while pending or running:
        while pending and len(running) < width:
            j = pending.pop(0)
            running.append((j, subprocess.Popen(j[1], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        j, p = running.pop(0)
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(" ".join(j[1]) + "\n" + out)
            raise RuntimeError("build failed: %s" % j[0])
        if verbose:
            sys.stderr.write("built %s\n" % j[0])


def _mirror_deps():
    deps = [os.path.join(SHIM, "t8mini.cpp"), os.path.join(SHIM, "t8.h"), os.path.join(HERE, "miniforest.c"),
            os.path.join(MIRROR_INC, "t8gpu_b200.h")]
    for d in (HARN, os.path.join(MIRROR_INC, "t8gpu")):
        for dp, _, fs in os.walk(d):
            deps += [os.path.join(dp, f) for f in fs]
    return deps


def build_mirror(force=False, verbose=False):
    """The reference's example solvers (unmodified, from /root/reference) over include/t8gpu/ -> libmirror_*.so.
    Also compiles the three main*.cu (objects only): all nine example TUs build against the mirror."""
    if not os.path.isdir(REF):
        return False
    if mirror_available() and not force:
        oldest = min(os.path.getmtime(os.path.join(OUT, "libmirror_%s_%s.so" % (k, p))) for k in LIBS for p in ("f32", "f64"))
        if oldest > max(os.path.getmtime(d) for d in _mirror_deps()):
            return True
    objdir = os.path.join(OUT, "obj_mirror")
    os.makedirs(objdir, exist_ok=True)
    jobs, by_lib = [], {}
    for prec in ("f32", "f64"):
        inc = ["-I", MIRROR_INC, "-I", SHIM] + (["-DT8GPU_FLOAT_TYPE=double"] if prec == "f64" else [])
        for kind, (tus, harness, exdir) in LIBS.items():
            objs = by_lib.setdefault((kind, prec), [])
            for tu in tus:
                obj = os.path.join(objdir, "%s_%s_%s.o" % (kind, prec, os.path.basename(tu).replace(".cu", "")))
                jobs.append((obj, ["nvcc"] + FLAGS + inc + ["-c", os.path.join(REF, tu), "-o", obj]))
                objs.append(obj)
            obj = os.path.join(objdir, "%s_%s_harness.o" % (kind, prec))
            jobs.append((obj, ["nvcc"] + FLAGS + inc + ["-DT8B200_MIRROR_BUILD", "-I", os.path.join(REF, exdir), "-c",
                                                        os.path.join(HARN, harness), "-o", obj]))
            objs.append(obj)
    inc = ["-I", MIRROR_INC, "-I", SHIM]
    for tu in EXAMPLE_TUS:
        if "main" in tu:
            obj = os.path.join(objdir, "main_%s.o" % tu.replace("/", "_").replace(".cu", ""))
            jobs.append((obj, ["nvcc"] + FLAGS + inc + ["-c", os.path.join(REF, tu), "-o", obj]))
    common = []
    for src, lang in ((os.path.join(SHIM, "t8mini.cpp"), "c++"), (os.path.join(HERE, "miniforest.c"), "c")):
        obj = os.path.join(objdir, os.path.basename(src).split(".")[0] + ".o")
        cmd = (["g++", "-std=c++17"] if lang == "c++" else ["gcc", "-std=gnu11"]) + ["-O2", "-fPIC", "-I", SHIM, "-c", src, "-o", obj]
        jobs.append((obj, cmd))
        common.append(obj)
    _run_jobs(jobs, verbose)
    for (kind, prec), objs in by_lib.items():
        so = os.path.join(OUT, "libmirror_%s_%s.so" % (kind, prec))
        subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs + common +
                              ["-L", PRODUCT_LIBDIR, "-lt8gpu_b200", "-Xlinker", "-rpath=" + PRODUCT_LIBDIR, "-Xlinker",
                               "-rpath=$ORIGIN/../../t8gpu_b200"])
    return True


def available():
    return all(os.path.exists(os.path.join(OUT, "libref_%s_%s.so" % (k, p))) for k in LIBS for p in ("f32", "f64"))


def _scratch_f64():
    d = os.path.join(OUT, "inc_f64", "t8gpu", "memory")
    os.makedirs(d, exist_ok=True)
    src = open(os.path.join(REF, "t8gpu/memory/memory_manager.h")).read()
    patched, n = re.subn(r"using float_type(\s*)= float;", r"using float_type\1= double;", src)
    assert n == 2, "expected exactly the two float_type lines of memory_manager.h:29,39"
    open(os.path.join(d, "memory_manager.h"), "w").write(patched)
    open(os.path.join(d, "memory_manager.inl"), "w").write(open(os.path.join(REF, "t8gpu/memory/memory_manager.inl")).read())
    return os.path.join(OUT, "inc_f64")


def build(force=False, verbose=False):
    if not os.path.isdir(REF):
        return False
    if available() and not force:
        newest = max(os.path.getmtime(os.path.join(dp, f)) for d in (SHIM, HARN) for dp, _, fs in os.walk(d) for f in fs)
        newest = max(newest, os.path.getmtime(os.path.join(HERE, "miniforest.c")))
        oldest = min(os.path.getmtime(os.path.join(OUT, "libref_%s_%s.so" % (k, p))) for k in LIBS for p in ("f32", "f64"))
        if oldest > newest:
            return True
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    inc64 = _scratch_f64()
    jobs = []
    for prec in ("f32", "f64"):
        inc = (["-I", inc64] if prec == "f64" else []) + ["-I", SHIM, "-I", REF]
        for kind, (tus, harness, exdir) in LIBS.items():
            for tu in tus:
                obj = os.path.join(OUT, "obj", "%s_%s_%s.o" % (kind, prec, os.path.basename(tu).replace(".cu", "")))
                jobs.append((kind, prec, obj, ["nvcc"] + FLAGS + inc + ["-c", os.path.join(REF, tu), "-o", obj]))
            obj = os.path.join(OUT, "obj", "%s_%s_harness.o" % (kind, prec))
            jobs.append((kind, prec, obj, ["nvcc"] + FLAGS + inc + ["-I", os.path.join(REF, exdir), "-c",
                                                                   os.path.join(HARN, harness), "-o", obj]))
    common = []
    for src, lang in ((os.path.join(SHIM, "t8mini.cpp"), "c++"), (os.path.join(HERE, "miniforest.c"), "c")):
        obj = os.path.join(OUT, "obj", os.path.basename(src).split(".")[0] + ".o")
        cmd = (["g++", "-std=c++17"] if lang == "c++" else ["gcc", "-std=gnu11"]) + ["-O2", "-fPIC", "-I", SHIM, "-c", src, "-o", obj]
        jobs.append(("common", "", obj, cmd))
        common.append(obj)
    # run at most 6 compiles at a time
    running, results = [], []
    pending = list(jobs)
    while pending or running:
        while pending and len(running) < 6:
            j = pending.pop(0)
            running.append((j, subprocess.Popen(j[3], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        j, p = running.pop(0)
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(" ".join(j[3]) + "\n" + out)
            raise RuntimeError("reference build failed: %s" % j[2])
        if verbose:
            sys.stderr.write("built %s\n" % j[2])
    for prec in ("f32", "f64"):
        for kind in LIBS:
            objs = [j[2] for j in jobs if j[0] == kind and j[1] == prec] + common
            so = os.path.join(OUT, "libref_%s_%s.so" % (kind, prec))
            subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs)
    return True


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_mirror(force="--force" in sys.argv, verbose=True))
