#!/usr/bin/env python
"""A/B builds of the library: tools/build_variant.py NAME [-DFLAG ...] compiles csrc/fused.cu, csrc/structured.cu and csrc/subgrid.cu with
the extra flags and links them with the other objects into t8gpu_b200/build/variants/libNAME.so (git-ignored, travels
with gpurun).  Select at run time with T8GPU_B200_LIB=<path>, so several variants are timed in ONE gpurun call."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from t8gpu_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    B.build()
    objdir = os.path.join(B.HERE, "build")
    vdir = os.path.join(objdir, "variants")
    os.makedirs(vdir, exist_ok=True)
    objs, procs = [], []
    for s in B.SOURCES:
        if s in ("fused.cu", "subgrid.cu", "structured.cu"):
            obj = os.path.join(vdir, name + "_" + s.replace(".cu", ".o"))
            cmd = [B._nvcc()] + B.NVCC_FLAGS + flags + ["-c", os.path.join(B.CSRC, s), "-o", obj]
            procs.append(subprocess.Popen(cmd))
        else:
            obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")
    lib = os.path.join(vdir, "lib%s.so" % name)
    subprocess.check_call([B._nvcc(), "-shared", "-o", lib] + objs +
                          ["-gencode", "arch=compute_100a,code=sm_100a", "-lpthread"])
    print(lib)


if __name__ == "__main__":
    main()
