"""Build recipe for libafb200.so (sm_100a only, in-tree so the .so travels to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libafb200.so")
SOURCES = ["af_host.cpp", "af_fastq.cpp", "af_kernels.cu", "af_pipeline.cu", "af_exchange.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-shared", "-cudart", "static"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "anchored_fusion.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> anchored_fusion_b200/libafb200.so"""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("AF_NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + ["-x", "cu"] + [os.path.join(CSRC, s) for s in SOURCES] + \
          ["-o", LIB, "-lz", "-lpthread"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports CC=/opt/gcc/bin/gcc, which lacks some spec files
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True, extra=sys.argv[1:])
