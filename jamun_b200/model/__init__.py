from . import arch
from .denoiser import Denoiser

__all__ = ["Denoiser", "arch"]
