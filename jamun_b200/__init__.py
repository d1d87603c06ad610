"""jamun_b200 -- B200-native walk-jump sampling hot path of prescient-design/jamun.

Public surface mirrors the reference's import paths for this path: ``jamun_b200.model.Denoiser``,
``jamun_b200.model.arch.E3Conv``, ``jamun_b200.e3tools.nn.{ConvBlock,Conv,EquivariantMLP}``,
``jamun_b200.sampling.{Sampler, walkjump.SingleMeasurementSampler, mcmc.{BAOAB,ABOBA}}``,
``jamun_b200.utils.ModelSamplingWrapper``, ``jamun_b200.distributions.ConstantSigma``.  ``import jamun`` (the alias
package at the repo root) resolves the reference's own ``jamun.*`` names to these modules.
"""
__version__ = "0.1.0"

from . import data, distributions, e3tools, irreps, lr_schedules, model, sampling, synthetic, utils  # noqa: E402,F401
from .factory import default_arch, default_denoiser  # noqa: E402,F401
