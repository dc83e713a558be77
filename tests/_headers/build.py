"""TEST INFRASTRUCTURE: builds the header-mirror checks.

  compile_check.cu   user-style translation unit against include/t8gpu/ (compile only)
  mesh_harness.cu    t8gpu::MeshManager<...,3> of include/t8gpu/mesh/mesh_manager.h over the t8mini stand-in for t8code
                     (oracle/ref_shim + oracle/miniforest.c), linked against libt8gpu_b200.so
                     -> libmeshharness_f32.so / libmeshharness_f64.so (git-ignored, travel to the GPU box)
  subgrid_harness.cu t8gpu::SubgridMeshManager<..., Subgrid<4,4,4>> and <..., Subgrid<4,4>> likewise
                     -> libsubgridharness_f32.so / libsubgridharness_f64.so
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
INC = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "ref_shim")]
FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr", "-O2", "-lineinfo",
         "-Xcompiler", "-fPIC", "-w"]
LIBDIR = os.path.join(ROOT, "t8gpu_b200")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    out = [os.path.join(HERE, "mesh_harness.cu"), os.path.join(HERE, "subgrid_harness.cu"),
           os.path.join(ROOT, "oracle", "miniforest.c"),
           os.path.join(ROOT, "oracle", "ref_shim", "t8mini.cpp"), os.path.join(ROOT, "oracle", "ref_shim", "t8.h"),
           os.path.join(ROOT, "include", "t8gpu_b200.h")]
    for dp, _, fs in os.walk(os.path.join(ROOT, "include", "t8gpu")):
        out += [os.path.join(dp, f) for f in fs]
    return out


def compile_check():
    """nvcc -c of the user-style TU for both precisions; raises on failure."""
    for extra in ([], ["-DT8GPU_FLOAT_TYPE=double"]):
        subprocess.check_call(["nvcc"] + FLAGS + INC + extra + ["-c", os.path.join(HERE, "compile_check.cu"), "-o",
                                                                os.devnull])


def build(force=False):
    objdir = os.path.join(HERE, "obj")
    os.makedirs(objdir, exist_ok=True)
    common = []
    for src, cc in ((os.path.join(ROOT, "oracle", "ref_shim", "t8mini.cpp"), ["g++", "-std=c++17"]),
                    (os.path.join(ROOT, "oracle", "miniforest.c"), ["gcc", "-std=gnu11"])):
        obj = os.path.join(objdir, os.path.basename(src).split(".")[0] + ".o")
        if force or _newer(obj, [src]):
            subprocess.check_call(cc + ["-O2", "-fPIC", "-I", os.path.join(ROOT, "oracle", "ref_shim"), "-c", src, "-o", obj])
        common.append(obj)
    procs = []
    for name, src in (("meshharness", "mesh_harness.cu"), ("subgridharness", "subgrid_harness.cu")):
        for prec, extra in (("f32", []), ("f64", ["-DT8GPU_FLOAT_TYPE=double"])):
            so = os.path.join(HERE, "lib%s_%s.so" % (name, prec))
            if not force and not _newer(so, _deps() + common):
                continue
            obj = os.path.join(objdir, "%s_%s.o" % (name, prec))
            p = subprocess.Popen(["nvcc"] + FLAGS + INC + extra + ["-c", os.path.join(HERE, src), "-o", obj])
            procs.append((p, so, obj))
    for p, so, obj in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed on %s" % obj)
        subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, obj] + common +
                              ["-L", LIBDIR, "-lt8gpu_b200", "-Xlinker", "-rpath=" + LIBDIR, "-Xlinker",
                               "-rpath=$ORIGIN/../../t8gpu_b200"])
    return True


if __name__ == "__main__":
    compile_check()
    print(build(force="--force" in sys.argv))
