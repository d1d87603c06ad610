"""BAOAB / ABOBA configuration dataclasses (mirror of /root/reference/src/jamun/sampling/mcmc/_splitting.py:12-58)."""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Callable, Optional, Union

import torch
from torch import Tensor

from .functional import aboba, baoab


def _check_v_init(v_init):
    if isinstance(v_init, str) and v_init not in {"gaussian", "zero"}:
        raise RuntimeError(f"{v_init} not in (gaussian, zero)")


@dataclass
class ABOBA:
    delta: float = 1.0
    friction: float = 1.0
    M: float = 1.0
    steps: int = 128
    save_trajectory: bool = False
    save_every_n_steps: int = 1
    burn_in_steps: int = 0
    verbose: bool = False
    cpu_offload: bool = False
    v_init: Union[str, Tensor] = "zero"
    inverse_temperature: float = 1.0
    score_fn_clip: Optional[float] = None

    def __post_init__(self):
        _check_v_init(self.v_init)

    def params(self, **kwargs):
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)} | kwargs

    def __call__(self, y: torch.Tensor, score_fn: Callable, **kwargs):
        return aboba(y, score_fn, **self.params(**kwargs))


@dataclass
class BAOAB:
    delta: float = 1.0
    friction: float = 1.0
    M: float = 1.0
    steps: int = 128
    save_trajectory: bool = False
    save_every_n_steps: int = 1
    burn_in_steps: int = 0
    verbose: bool = False
    cpu_offload: bool = False
    v_init: Union[str, Tensor] = "zero"
    inverse_temperature: float = 1.0
    score_fn_clip: Optional[float] = None

    def __post_init__(self):
        _check_v_init(self.v_init)

    def params(self, **kwargs):
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)} | kwargs

    def __call__(self, y: torch.Tensor, score_fn: Callable, **kwargs):
        return baoab(y, score_fn, **self.params(**kwargs))
