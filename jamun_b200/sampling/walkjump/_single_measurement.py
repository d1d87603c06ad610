"""Single-measurement walk-jump sampler (mirror of
/root/reference/src/jamun/sampling/walkjump/_single_measurement.py:8-89)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from ..mcmc import BAOAB


class SingleMeasurementSampler:
    """Single Measurement Walk-Jump Sampler."""

    def __init__(self, mcmc, sigma: float, y_init_distribution: Optional[torch.distributions.Distribution] = None):
        self.mcmc = mcmc
        self.sigma = float(sigma)
        self.y_init_distribution = y_init_distribution

    def _y_init(self, model, batch_size, y_init):
        if y_init is None:
            if self.y_init_distribution is None:
                raise RuntimeError("either y_init and y_init_distribution must be supplied")
            y_init = self.y_init_distribution.sample(sample_shape=(batch_size,)).to(model.device)
        return y_init

    def walk(self, model, batch_size: Optional[int] = None, y_init: Optional[torch.Tensor] = None,
             v_init: str | Tensor = "gaussian", **kwargs):
        y_init = self._y_init(model, batch_size, y_init)
        if isinstance(self.mcmc, BAOAB) and hasattr(model, "fused_walk"):
            out = model.fused_walk(self.mcmc, y_init, self.sigma, v_init=v_init, **kwargs)  # hot path
            y_traj = out["y_traj"]
        else:
            y, v, y_traj, score_traj = self.mcmc(y_init, lambda y: model.score(y, self.sigma), v_init=v_init, **kwargs)
            out = {"y": y, "v": v, "y_traj": y_traj, "score_traj": score_traj}
        out["t_traj"] = torch.ones(y_traj.size(0), device=y_traj.device, dtype=int) if y_traj is not None else None
        return out

    def walk_jump(self, model, batch_size: Optional[int] = None, y_init: Optional[torch.Tensor] = None,
                  v_init: str | Tensor = "gaussian", **kwargs):
        out = self.walk(model, batch_size=batch_size, y_init=y_init, v_init=v_init, **kwargs)
        y, y_traj = out["y"], out["y_traj"]
        if "xhat" not in out:  # generic protocol: jump needs its own evaluations (ABOBA / foreign models)
            out["xhat"] = model.xhat(y, sigma=self.sigma)
            if y_traj is not None:
                out["xhat_traj"] = torch.stack(
                    [model.xhat(y_traj[i, :].to(model.device), sigma=self.sigma) for i in range(y_traj.size(0))], dim=0)
            else:
                out["xhat_traj"] = None
        return {k: out.get(k) for k in ("xhat", "y", "v", "xhat_traj", "y_traj", "t_traj", "score_traj")}

    def sample(self, model, batch_size: Optional[int] = None, y_init: Optional[torch.Tensor] = None,
               v_init: str | Tensor = "gaussian", **kwargs):
        out = self.walk_jump(model, batch_size=batch_size, y_init=y_init, v_init=v_init, **kwargs)
        out["sample"] = out["xhat"]
        return out
