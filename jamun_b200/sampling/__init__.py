from . import mcmc, walkjump
from ._sampler import Sampler

__all__ = ["Sampler", "mcmc", "walkjump"]
