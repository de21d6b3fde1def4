"""The driver's smoke entry point (`__graft_entry__.smoke()`: a few fused training steps on a minimal scene, then march / encoder /
compositing checked against the oracle) must keep working as the kernels change."""
import pytest

pytestmark = pytest.mark.gpu


def test_graft_entry_smoke(cuda_dev):
    import __graft_entry__ as g
    g.smoke()
