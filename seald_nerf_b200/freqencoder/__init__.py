from .freq import FreqEncoder, freq_encode

__all__ = ["FreqEncoder", "freq_encode"]
