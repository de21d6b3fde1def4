"""Recipe: compile the reference's own CUDA extensions into ``oracle/_ref/`` (TEST INFRASTRUCTURE ONLY).

The sources are compiled *where they lie* under ``/root/reference`` (nothing is copied into this
repo); only the built ``.so`` files land in ``oracle/_ref/`` (git-ignored, shipped to the GPU box by
gpurun).  They are the executable oracle the ``-m gpu`` parity tests compare our sm_100a kernels
against.  The only deviation from the reference build files
(``/root/reference/gridencoder/backend.py:6-9`` etc.) is ``-std=c++17`` (torch 2.11 needs it).

Usage:  python oracle/build_ref.py [gridencoder raymarching freqencoder shencoder ffmlp]
Each extension is built in its own process (they take 3-8 min each; they run in parallel).
"""
import glob
import os
import subprocess
import sys

REF = os.environ.get("SEALD_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

EXTS = {
    # name -> (module name, sources, extra nvcc flags, extra include paths)
    "gridencoder": ("_ref_gridencoder", ["gridencoder.cu", "bindings.cpp"], [], []),
    "raymarching": ("_ref_raymarching", ["raymarching.cu", "bindings.cpp"], [], []),
    "freqencoder": ("_ref_freqencoder", ["freqencoder.cu", "bindings.cpp"], [], []),
    "shencoder": ("_ref_shencoder", ["shencoder.cu", "bindings.cpp"], [], []),
    "ffmlp": ("_ref_ffmlp", ["ffmlp.cu", "bindings.cpp"],
              ["--expt-extended-lambda", "--expt-relaxed-constexpr", "-Xcompiler=-mf16c",
               "-Xcompiler=-Wno-float-conversion", "-Xcompiler=-fno-strict-aliasing"],
              "CUTLASS"),
}


def _cutlass_includes():
    import site
    for sp in site.getsitepackages():
        base = os.path.join(sp, "flashinfer", "data", "cutlass")
        if os.path.isdir(os.path.join(base, "include")):
            return [os.path.join(base, "include"), os.path.join(base, "tools", "util", "include")]
    return []


def build_one(name):
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "2")
    from torch.utils.cpp_extension import load
    mod, srcs, extra, inc = EXTS[name]
    bdir = os.path.join(OUT, "build_" + name)
    os.makedirs(bdir, exist_ok=True)
    common = ["-O3", "-std=c++17"]
    nv = common + ["-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                   "-U__CUDA_NO_HALF2_OPERATORS__"] + extra
    load(name=mod, extra_cflags=common, extra_cuda_cflags=nv,
         extra_include_paths=_cutlass_includes() if inc == "CUTLASS" else [],
         sources=[os.path.join(REF, name, "src", s) for s in srcs],
         build_directory=bdir, verbose=False, is_python_module=False)
    so = glob.glob(os.path.join(bdir, mod + "*.so"))[0]
    dst = os.path.join(OUT, mod + ".so")
    if os.path.exists(dst):
        os.remove(dst)
    os.replace(so, dst)
    print("built", dst, flush=True)


def have(name):
    return os.path.exists(os.path.join(OUT, EXTS[name][0] + ".so"))


def main(argv):
    if len(argv) >= 2 and argv[0] == "--one":
        build_one(argv[1])
        return 0
    names = argv or ["gridencoder", "raymarching", "freqencoder", "shencoder", "ffmlp"]
    if not os.path.isdir(REF):
        print("reference tree absent (%s): using prebuilt oracle/_ref/*.so as-is" % REF)
        return 0
    os.makedirs(OUT, exist_ok=True)
    # the reference's host modules, byte-compiled where they lie into oracle/_ref/py/ (sourceless .pyc; oracle/ref_runtime.py)
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_runtime
    ref_runtime.stage_python()
    procs = []
    for n in names:
        if have(n):
            continue
        log = open(os.path.join(OUT, "build_%s.log" % n), "w")
        procs.append((n, subprocess.Popen([sys.executable, __file__, "--one", n], stdout=log, stderr=log)))
    rc = 0
    for n, p in procs:
        r = p.wait()
        print(n, "rc", r, flush=True)
        rc |= r
    return rc


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
