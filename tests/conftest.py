import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_ref(name):
    """Import one of the reference's own extensions built by oracle/build_ref.py (oracle/_ref/_ref_<name>.so).
    Returns None when it is not there (the tests then compare against the CPU oracle only)."""
    import torch  # noqa: F401  (the .so links against torch)
    path = os.path.join(ROOT, "oracle", "_ref", "_ref_%s.so" % name)
    if not os.path.exists(path):
        return None
    modname = "_ref_%s" % name
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[modname] = mod
    return mod


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


GOLDEN = os.path.join(ROOT, "tests", "golden")
