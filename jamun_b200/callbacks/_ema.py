"""Exponential moving average of the parameters (mirror of /root/reference/src/jamun/callbacks/_ema.py: EMA / EMAOptimizer
semantics -- ema <- decay*ema + (1-decay)*p every `every_n_steps` optimiser steps, swap in for evaluation).  The update is one
kernel per flat parameter buffer (``jamun_ema_update``)."""
from __future__ import annotations

import contextlib
from typing import List

import torch

from .. import ops


class EMA:
    def __init__(self, decay: float, validate_original_weights: bool = False, every_n_steps: int = 1, cpu_offload: bool = False):
        if not (0 <= decay <= 1):
            raise ValueError("EMA decay value must be between 0 and 1")
        self.decay, self.every_n_steps = decay, every_n_steps
        self.validate_original_weights, self.cpu_offload = validate_original_weights, cpu_offload
        self.step = 0
        self._params: List[torch.nn.Parameter] = []
        self._ema: List[torch.Tensor] = []

    def on_fit_start(self, module: torch.nn.Module) -> None:
        self._params = [p for p in module.parameters() if p.requires_grad]
        self._ema = [p.detach().clone().contiguous() for p in self._params]

    def on_train_batch_end(self) -> None:
        """Call after optimizer.step()."""
        self.step += 1
        if self.step % self.every_n_steps:
            return
        for e, p in zip(self._ema, self._params):
            if not p.is_cuda:
                raise RuntimeError("jamun_b200 EMA updates CUDA parameters only (no CPU fallback)")
            ops.ema_update(e, p.detach().contiguous(), self.decay)

    def swap_model_weights(self) -> None:
        for e, p in zip(self._ema, self._params):
            tmp = p.detach().clone()
            p.data.copy_(e)
            e.copy_(tmp)

    @contextlib.contextmanager
    def swapped(self):
        self.swap_model_weights()
        try:
            yield
        finally:
            self.swap_model_weights()

    def state_dict(self):
        return {"step": self.step, "ema": [e.cpu() for e in self._ema]}
