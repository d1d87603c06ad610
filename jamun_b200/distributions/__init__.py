"""Noise-level (sigma) distributions of the Denoiser's `sigma_distribution` argument and of the measurement samplers
(mirror of /root/reference/src/jamun/distributions/_distributions.py:5-111: same class names, constructor arguments and
sampling arithmetic, so the reference's Hydra targets `jamun.distributions.*` resolve).  Host-side torch only: one scalar per
training step."""
import torch


class ConstantSigma(torch.distributions.Distribution):
    def __init__(self, sigma: float):
        self.sigma = torch.tensor(sigma)

    def sample(self, sample_shape: torch.Size = torch.Size([])):
        return self.sigma.expand(sample_shape)

    def __repr__(self):
        return f"Constant(sigma={self.sigma})"


class UniformSigma(torch.distributions.Uniform):
    def __init__(self, sigma_max, sigma_min=1e-4):
        self.sigma_max, self.sigma_min = sigma_max, sigma_min
        super().__init__(low=sigma_min, high=sigma_max)

    def __repr__(self):
        return f"UniformSigma(sigma_max={self.sigma_max}, sigma_min={self.sigma_min})"


class ExponentialSigma(torch.distributions.Distribution):
    """sigma_min * (sigma_max / sigma_min)^t with t ~ U(epsilon, 1): log-uniform noise levels."""

    def __init__(self, sigma_max=50.0, sigma_min=1e-2, epsilon=1e-5):
        self.sigma_max, self.sigma_min, self.epsilon = sigma_max, sigma_min, epsilon
        self.t_dist = torch.distributions.Uniform(epsilon, 1.0)

    def sample(self, sample_shape=torch.Size([])):
        t = self.t_dist.sample(sample_shape)
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def __repr__(self):
        return f"ExponentialSigma(sigma_max={self.sigma_max}, sigma_min={self.sigma_min})"


class ClippedLogNormalSigma(torch.distributions.Distribution):
    def __init__(self, log_sigma_mean: float, log_sigma_std: float, sigma_max: float = 100.0):
        self.log_sigma_mean, self.log_sigma_std = log_sigma_mean, log_sigma_std
        self.log_sigma_dist = torch.distributions.Normal(log_sigma_mean, log_sigma_std)
        self.sigma_max = sigma_max

    def sample(self, sample_shape=torch.Size([])):
        return torch.clamp(self.log_sigma_dist.sample(sample_shape).exp(), max=self.sigma_max)

    def __repr__(self):
        return f"LogNormalSigma(log_sigma_mean={self.log_sigma_mean}, log_sigma_std={self.log_sigma_std}, sigma_max={self.sigma_max})"


class UniformPlusNormal(torch.distributions.Distribution):
    """U(0, 1) + sigma * N(0, 1), element-wise over sample_shape (the toy `y_init_distribution` of the samplers)."""

    def __init__(self, sigma, sample_shape, dtype=None):
        self.sigma, self.sample_shape, self.dtype = sigma, sample_shape, dtype

    def sample(self, sample_shape=torch.Size([])):
        unit = torch.distributions.Uniform(torch.tensor(0.0, dtype=self.dtype), torch.tensor(1.0, dtype=self.dtype))
        x = unit.sample((*sample_shape, *self.sample_shape))
        return x + torch.randn_like(x) * self.sigma

    def __repr__(self):
        return f"UniformPlusNormal(sigma={self.sigma}, sample_shape={self.sample_shape})"


class CategoricalValue(torch.distributions.Distribution):
    """A categorical distribution over a table of values."""

    def __init__(self, values: torch.Tensor, categorical: torch.distributions.Categorical):
        if values.shape[0] != categorical.probs.shape[0]:
            raise RuntimeError(f"{values.shape[0]=} != {categorical.probs.shape[0]=}")
        self.values, self.categorical = values, categorical

    def sample(self, sample_shape=torch.Size([])):
        return self.values[self.categorical.sample(sample_shape)]

    @property
    def mean(self):
        probs = self.categorical.probs
        return (self.values * probs.reshape(probs.shape + (1,) * (self.values.ndim - probs.ndim))).sum(0)

    def __repr__(self):
        return "CategoricalValue"


class WeightedMeasurement(CategoricalValue):
    """Noise level of the k-th of m averaged measurements, sigma / sqrt(k), drawn with the given probabilities."""

    def __init__(self, sigma: float, probs: torch.Tensor):
        self.sigma, self.m = sigma, probs.shape[0]
        super().__init__(values=sigma * torch.arange(1, self.m + 1).pow(-0.5), categorical=torch.distributions.Categorical(probs=probs))

    def __repr__(self):
        return f"WeightedMeasurement(sigma={self.sigma}, m={self.m})"


class UniformMeasurement(WeightedMeasurement):
    def __init__(self, sigma: float, m: int):
        super().__init__(sigma=sigma, probs=torch.ones(m))

    def __repr__(self):
        return f"UniformMeasurement(sigma={self.sigma}, m={self.m})"


__all__ = ["CategoricalValue", "ClippedLogNormalSigma", "ConstantSigma", "ExponentialSigma", "UniformMeasurement", "UniformPlusNormal",
           "UniformSigma", "WeightedMeasurement"]
