from .ffmlp import FFMLP, ffmlp_forward

__all__ = ["FFMLP", "ffmlp_forward"]
