"""Builds roitr_b200/lib/libroitr_b200.so from roitr_b200/csrc/*.cu with nvcc for sm_100a (in-tree, so the .so
travels to the GPU box with the gpurun snapshot). No torch headers: every translation unit compiles in seconds."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
SO = os.path.join(LIBDIR, "libroitr_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
         "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build_variant(name, defines):
    """A second library with extra -D flags (scripts/gpu_ab.sh: A/B runs of compile-time choices via ROITR_B200_LIB)."""
    vdir = os.path.join(LIBDIR, "variants")
    odir = os.path.join(vdir, "obj_" + name)
    os.makedirs(odir, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))

    def one(s):
        o = os.path.join(odir, os.path.basename(s)[:-3] + ".o")
        subprocess.check_call([NVCC] + ARCH + FLAGS + ["-D" + d for d in defines] + ["-c", s, "-o", o])
        return o
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, srcs))
    so = os.path.join(vdir, "libroitr_b200_%s.so" % name)
    subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", so] + objs + ["-lcudart"])
    return so


def build(force=False, verbose=False, ptxas_info=False):
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    jobs = []
    for s in srcs:
        o = os.path.join(OBJDIR, os.path.basename(s)[:-3] + ".o")
        if force or not _newer(o, [s] + hdrs):
            cmd = [NVCC] + ARCH + FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", s, "-o", o]
            jobs.append((s, cmd))

    def run(job):
        s, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    failed = False
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(run, jobs):
            if r.returncode != 0:
                failed = True
                sys.stderr.write("nvcc failed for %s\n%s\n%s\n" % (s, r.stdout, r.stderr))
            elif verbose or ptxas_info:
                sys.stderr.write("== %s\n%s%s" % (os.path.basename(s), r.stdout, r.stderr))
    if failed:
        raise RuntimeError("nvcc failed")
    objs = [os.path.join(OBJDIR, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if jobs or not os.path.exists(SO):
        subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", SO] + objs + ["-lcudart"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv))
