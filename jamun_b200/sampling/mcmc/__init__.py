from . import functional
from ._splitting import ABOBA, BAOAB

__all__ = ["ABOBA", "BAOAB", "functional"]
