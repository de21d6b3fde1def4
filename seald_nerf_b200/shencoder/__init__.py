from .sphere_harmonics import SHEncoder, sh_encode

__all__ = ["SHEncoder", "sh_encode"]
