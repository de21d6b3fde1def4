"""Build libseald_b200.so (hand-written sm_100a kernels + C-ABI) in-tree with plain nvcc.

    python -m seald_nerf_b200.build [--force] [--verbose]

No torch dependency: the library takes raw device pointers (include/seald_b200.h).  Objects are rebuilt
only when a source/header is newer.  The march/composite kernels rely on IEEE arithmetic with nvcc's default
FMA contraction (bit-exact parity with the reference), so --use_fast_math must never be added here.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libseald_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-DSEALD_SM_ARCH=100",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def _compile(src, obj, verbose):
    cmd = [NVCC] + ARCH + FLAGS + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    log = obj + ".log"
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, p.stderr[-6000:]))
    if verbose:
        print(p.stderr)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    hdr_mtime = max(os.path.getmtime(h) for h in hdrs)
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(s, o) or hdr_mtime > os.path.getmtime(o):
            jobs.append((s, o))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            futs = [ex.submit(_compile, s, o, verbose) for s, o in jobs]
            for f in futs:
                f.result()
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
