"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            txt = open(os.path.join(ROOT, "include", fn)).read()
            names += re.findall(r"^\s*(?:int|uint64_t|const char\*)\s+(seald_\w+)\s*\(", txt, flags=re.M)
    return sorted(set(names))


def test_header_symbols_exported():
    from seald_nerf_b200 import _lib
    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libseald_b200.so does not export %s" % name
    # and the Python binding table covers exactly the header
    assert sorted(_lib.exported_symbols()) == declared


def test_version_and_arch():
    from seald_nerf_b200 import _lib
    lib = _lib.load()
    assert lib.seald_version() >= 100
    assert lib.seald_sm_arch() == 100
    assert lib.seald_strerror(-2).decode().startswith("unsupported")


def test_library_has_sm100a_code_only():
    from seald_nerf_b200 import _lib
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from seald_nerf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except RuntimeError as e:
        assert "no CPU" in str(e) or "not found" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")


def test_aliases_install():
    import seald_nerf_b200
    seald_nerf_b200.install_aliases()
    import raymarching  # noqa: F401
    from gridencoder import GridEncoder  # noqa: F401
    from encoding import get_encoder
    enc, dim = get_encoder("hashgrid", desired_resolution=2048)
    assert dim == 32 and tuple(enc.embeddings.shape) == (6119864, 2)
    enc, dim = get_encoder("frequency", multires=10)
    assert dim == 63
