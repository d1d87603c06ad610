"""Sigma distributions (mirror of /root/reference/src/jamun/distributions/_distributions.py:42-108)."""
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


class ClippedLogNormalSigma(torch.distributions.Distribution):
    def __init__(self, log_sigma_mean: float, log_sigma_std: float, sigma_max: float = 100.0):
        self.log_sigma_dist = torch.distributions.Normal(log_sigma_mean, log_sigma_std)
        self.sigma_max = sigma_max

    def sample(self, sample_shape=torch.Size([])):
        return torch.clamp(self.log_sigma_dist.sample(sample_shape).exp(), max=self.sigma_max)
