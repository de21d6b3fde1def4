"""Drop-in `SealDNeRF.network.NeRFNetwork` (reference: SealDNeRF/network.py:10-279): the same D-NeRF field as
`dnerf.network.NeRFNetwork` (the reference file is a copy of it) on top of `SealNeRFTeacherRenderer`."""
from ..dnerf.network import NeRFNetwork as _DNeRFNetwork
from .renderer import SealNeRFTeacherRenderer


class NeRFNetwork(_DNeRFNetwork, SealNeRFTeacherRenderer):
    """MRO: field methods from dnerf.network.NeRFNetwork, `run_cuda` / mapper handling from SealNeRFTeacherRenderer."""
    pass
