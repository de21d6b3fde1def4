"""TEST INFRASTRUCTURE ONLY — the REFERENCE's own Python host code (dnerf/network.py, dnerf/renderer.py, SealDNeRF/renderer.py,
raymarching/raymarching.py, gridencoder/grid.py, ffmlp/ffmlp.py, ...) made importable beside this package, backed by the
reference's own CUDA extensions built unmodified into oracle/_ref/_ref_*.so (oracle/build_ref.py).

The reference is a script tree, not an installable package, and /root/reference does not exist on the GPU box.  So — exactly like
the `.cu` sources, which are compiled where they lie and ship only as `.so` — the Python modules the hot path needs are
byte-compiled where they lie (`stage_python()`, run by oracle/build_ref.py in the build container) and ship only as marshalled
code objects (`*.rbc`: a .pyc under another suffix, because the gpurun snapshot drops `*.pyc`) under oracle/_ref/py/ (git-ignored,
not gpurun-ignored), loaded by the small meta-path finder below.  No reference source text enters the repository.

`install()` puts that directory on sys.path, registers the five extensions under the names the reference's wrappers import
(`_raymarching`, `_gridencoder`, `_freqencoder`, `_shencoder`, `_ffmlp`) and stubs the third-party modules that only the
reference's GUI / mesh tooling touches.  The product package is never aliased in this mode: `import raymarching` is the
REFERENCE's wrapper, `seald_nerf_b200.raymarching` is ours, and the parity tests / `ref_gpu` baseline call both in one process.

Only tests/, scripts run by them, `__graft_entry__.smoke()` and bench.py's reference legs may import this module.
"""
import importlib.abc
import importlib.machinery
import importlib.util
import marshal
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
PYDIR = os.path.join(OUT, "py")
REF = os.environ.get("SEALD_REFERENCE_ROOT", "/root/reference")

# the modules on (or directly beside) the hot path, SURVEY.md §8(a)/(b)
PY_FILES = [
    "activation.py", "encoding.py",
    "raymarching/__init__.py", "raymarching/raymarching.py",
    "gridencoder/__init__.py", "gridencoder/grid.py",
    "freqencoder/__init__.py", "freqencoder/freq.py",
    "shencoder/__init__.py", "shencoder/sphere_harmonics.py",
    "ffmlp/__init__.py", "ffmlp/ffmlp.py",
    "nerf/__init__.py", "nerf/renderer.py", "nerf/network.py", "nerf/network_ff.py",
    "dnerf/renderer.py", "dnerf/network.py",
    "SealDNeRF/renderer.py", "SealDNeRF/network.py",
    "SealNeRF/renderer.py", "SealNeRF/network.py", "SealNeRF/seal_utils.py", "SealNeRF/color_utils.py", "SealNeRF/types.py",
]
EXTENSIONS = {"raymarching": "_raymarching", "gridencoder": "_gridencoder", "freqencoder": "_freqencoder", "shencoder": "_shencoder",
              "ffmlp": "_ffmlp"}
_STUBS = ("trimesh", "trimesh.creation", "trimesh.primitives", "json5", "pytorch3d", "pytorch3d.structures", "skspatial",
          "skspatial.objects", "open3d", "mcubes", "tensorboardX", "lpips", "torch_ema", "torchmetrics", "torchmetrics.functional",
          "imageio", "dearpygui", "dearpygui.dearpygui", "matplotlib", "matplotlib.pyplot", "turtle", "cv2", "sklearn.neighbors",
          "sklearn.decomposition")


def stage_python(force=False):
    """Byte-compile the reference modules of PY_FILES from /root/reference into oracle/_ref/py/ (sourceless .pyc)."""
    if not os.path.isdir(REF):
        return False
    n = 0
    for rel in PY_FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(PYDIR, rel[:-3] + ".rbc")
        if not os.path.exists(src):
            continue
        if force or not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            n += 1
    return True


def available(exts=("raymarching", "gridencoder", "freqencoder", "shencoder")):
    return os.path.exists(os.path.join(PYDIR, "dnerf", "network.rbc")) and all(
        os.path.exists(os.path.join(OUT, "_ref_%s.so" % e)) for e in exts)


class _RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports `a.b` from oracle/_ref/py/a/b.rbc (module), a/b/__init__.rbc (package) or the directory a/b/ (namespace package, like
    the reference's dnerf/ and SealDNeRF/ which have no __init__.py)."""

    def find_spec(self, fullname, path=None, target=None):
        base = os.path.join(PYDIR, *fullname.split("."))
        if os.path.exists(base + ".rbc"):
            return importlib.machinery.ModuleSpec(fullname, self, origin=base + ".rbc")
        if os.path.isdir(base):
            init = os.path.join(base, "__init__.rbc")
            spec = importlib.machinery.ModuleSpec(fullname, self, origin=init if os.path.exists(init) else base, is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        origin = module.__spec__.origin
        if not origin.endswith(".rbc"):
            return  # namespace package
        module.__file__ = origin
        with open(origin, "rb") as f:
            code = marshal.loads(f.read()[16:])  # 16-byte .pyc header (magic, flags, hash)
        exec(code, module.__dict__)


class _Any:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


def _stub_getattr(k):
    if k.startswith("__"):  # (inspect / importlib probe __file__, __path__, __wrapped__ ... on every module in sys.modules)
        raise AttributeError(k)
    return _Any()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__dict__.setdefault("__getattr__", _stub_getattr)
    sys.modules[name] = m
    return m


def load_extension(name):
    """oracle/_ref/_ref_<name>.so (PyInit__ref_<name>) registered under the module name the reference's wrapper imports."""
    import torch  # noqa: F401  (the .so links against libtorch)
    modname = "_ref_%s" % name
    mod = sys.modules.get(modname)
    if mod is None:
        path = os.path.join(OUT, modname + ".so")
        if not os.path.exists(path):
            return None
        spec = importlib.util.spec_from_file_location(modname, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules[modname] = mod
    sys.modules[EXTENSIONS[name]] = mod
    return mod


_installed = False


def install():
    """Make the reference's host modules importable in this process.  Raises if they were not staged / built."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference runtime not staged: run `python oracle/build_ref.py` where /root/reference exists")
    import torch
    for name in ("raymarching", "gridencoder", "freqencoder", "shencoder", "ffmlp"):
        if name in sys.modules and "seald_nerf_b200" in getattr(sys.modules[name], "__name__", ""):
            raise RuntimeError("seald_nerf_b200.install_aliases() is active: the reference runtime needs those names")
        load_extension(name)
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)

    def custom_meshgrid(*args):  # nerf/utils.py:41-46 (torch >= 1.10 branch); the real module drags in the whole trainer stack
        return torch.meshgrid(*args, indexing="ij")

    for name in ("nerf.utils", "dnerf.utils", "SealDNeRF.utils", "SealNeRF.utils"):
        _stub(name, custom_meshgrid=custom_meshgrid, Trainer=object)
    sys.meta_path.insert(0, _RefFinder())
    _installed = True


def dnerf_network(seald=False, **kw):
    """The reference's D-NeRF field + renderer (dnerf/network.py:10 or its SealD copy SealDNeRF/network.py:10) with the
    benchmark's configuration (BASELINE configs[1])."""
    install()
    if seald:
        from SealDNeRF.network import NeRFNetwork
    else:
        from dnerf.network import NeRFNetwork
    args = dict(encoding="hashgrid", bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=10)
    args.update(kw)
    return NeRFNetwork(**args)


if __name__ == "__main__":
    print("staged" if stage_python(force="--force" in sys.argv) else "reference tree absent", PYDIR)
