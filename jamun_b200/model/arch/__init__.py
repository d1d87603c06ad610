from .e3conv import E3Conv

__all__ = ["E3Conv"]
