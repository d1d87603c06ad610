"""Learning-rate multipliers for torch.optim.lr_scheduler.LambdaLR (the `lr_scheduler_config` of the Denoiser;
mirror of /root/reference/src/jamun/lr_schedules/_lr_schedules.py:2-21 -- same names and keyword-only arguments, so the
reference's Hydra configs resolve).  Host-side scalars only."""
from __future__ import annotations


def linear_warmup_linear_decay_lr_lambda(current_step: int, *, num_warmup_steps: int, num_training_steps: int) -> float:
    """0 -> 1 over the warm-up, then 1 -> 0 at num_training_steps (clamped at 0 afterwards)."""
    if current_step < num_warmup_steps:
        return current_step / max(1, num_warmup_steps)
    remaining = num_training_steps - current_step
    return max(0.0, remaining / max(1, num_training_steps - num_warmup_steps))


def linear_warmup_plateau_lr_lambda(current_step: int, *, num_warmup_steps: int, start_factor: float = 0.0,
                                    end_factor: float = 1.0) -> float:
    """start_factor -> end_factor over the warm-up, end_factor afterwards."""
    f = min(1.0, current_step / num_warmup_steps)
    return start_factor * (1 - f) + f * end_factor


def linear(current_step: int, *, start_factor: float = 0.0, slope: float = 1e-6) -> float:
    return max(0.0, start_factor + current_step * slope)


__all__ = ["linear", "linear_warmup_linear_decay_lr_lambda", "linear_warmup_plateau_lr_lambda"]
