"""Build recipe for libafb200.so (sm_100a only, in-tree so the .so travels to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libafb200.so")
SOURCES = ["af_host.cpp", "af_fastq.cpp", "af_kernels.cu", "af_tail.cu", "af_pipeline.cu", "af_exchange.cu", "af_genome.cu", "af_genome_host.cpp"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-shared", "-cudart", "static"]


OBJDIR = os.path.join(HERE, "build")
COMMON_DEPS = [os.path.join(CSRC, "af_common.h"), os.path.join(CSRC, "af_device.cuh"), os.path.join(CSRC, "af_inflate.h"), os.path.join(CSRC, "af_inflate_par.h"), os.path.join(CSRC, "af_crc32.h"), os.path.join(HERE, "..", "include", "anchored_fusion.h"), os.path.abspath(__file__)]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> anchored_fusion_b200/libafb200.so
    One object per source (rebuilt only when it or a header changed), then one link."""
    nvcc = os.environ.get("AF_NVCC", "nvcc")
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports CC=/opt/gcc/bin/gcc, which lacks some spec files
    env.pop("CXX", None)
    os.makedirs(OBJDIR, exist_ok=True)
    tag = os.path.join(OBJDIR, "flags.txt")
    flags = " ".join(NVCC_FLAGS + list(extra))
    if not os.path.exists(tag) or open(tag).read() != flags:
        force = True
    objs, jobs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        path = os.path.join(CSRC, src)
        if force or _newer(obj, [path] + COMMON_DEPS):
            jobs.append([nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + list(extra) + ["-x", "cu", "-c", path, "-o", obj])
    if jobs:
        from concurrent.futures import ThreadPoolExecutor

        def run(cmd):
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            subprocess.check_call(cmd, env=env)

        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:     # the translation units compile side by side
            list(pool.map(run, jobs))
    rebuilt = bool(jobs)
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB, "-lz", "-lpthread"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd, env=env)
        with open(tag, "w") as fh:
            fh.write(flags)
    return LIB


if __name__ == "__main__":
    args = sys.argv[1:]
    build(force="--force" in args, verbose=True, extra=[a for a in args if a != "--force"])
