"""Langevin splitting integrators (mirror of /root/reference/src/jamun/sampling/mcmc/functional/_splitting.py).

Two ways in:
* ``baoab`` / ``aboba`` keep the reference's protocol -- any callable ``score_fn(y) -> [N,3]`` -- and run the
  per-step update in the fused K6 kernels (clip, B/A/O/A, Philox draw);
* ``fused_baoab`` is the hot path used by SingleMeasurementSampler when the model is a jamun_b200 Denoiser:
  denoiser evaluation, score, jump and integrator update are one kernel sequence per step with no host
  synchronisation, and the jump ``xhat`` of every saved frame comes for free.

The reference's quirks are preserved: ``exp(-friction)`` without delta, the closing B half-kick without ``u``,
``range(1, steps)`` (steps-1 updates), ``score_traj`` always holding the initial score.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, Optional, Tuple, Union

import torch

from .... import _lib, engine, ops

_call_counter = itertools.count(1)


def _philox_seed() -> int:
    return int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF


def initialize_velocity(v_init: Union[str, torch.Tensor], y: torch.Tensor, u: float, noise: Optional[torch.Tensor] = None,
                        stream_id: int = 0) -> torch.Tensor:
    """Initialize velocity according to the given method."""
    if isinstance(v_init, str):
        if v_init == "gaussian":
            out = torch.empty_like(y)
            return ops.gaussian_axpy(None, 0.0, math.sqrt(u), noise, _philox_seed(), (stream_id << 32), out)
        if v_init == "zero":
            return torch.zeros_like(y)
        raise RuntimeError(f"{v_init} not in (gaussian, zero)")
    if isinstance(v_init, torch.Tensor):
        return v_init.clone()
    raise RuntimeError(f"{type(v_init)=} must be either `str` or `Tensor`.")


def create_score_fn(score_fn: Callable, inverse_temperature: float, score_fn_clip: Optional[float]) -> Callable:
    """Reference-compatible helper (torch ops; the integrators below clip inside the kernels instead)."""

    def score_fn_processed(y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        orig_score = score_fn(y).to(dtype=y.dtype)
        score = orig_score
        if score_fn_clip is not None:
            norm = torch.linalg.vector_norm(score, dim=-1, keepdim=True)
            clip = torch.min(norm, torch.ones_like(norm) * score_fn_clip)
            score = (score / norm) * clip
        return score * inverse_temperature, orig_score

    return score_fn_processed


def _walk_params(delta, friction, M, inverse_temperature, score_fn_clip, **extra) -> "_lib.WalkParams":
    u = pow(M, -1)
    return _lib.WalkParams(delta=float(delta), u=float(u), a=math.exp(-friction),
                           z_sqrt_u=math.sqrt(1 - math.exp(-2 * friction)) * math.sqrt(u),
                           beta=float(inverse_temperature), clip=float(score_fn_clip) if score_fn_clip else 0.0,
                           c_in=1.0, c_skip=0.0, c_out=0.0, sigma2=1.0, first=0, last=0, center=0, seed=_philox_seed(),
                           step=0, **extra)


def _saves(i: int, save_trajectory: bool, save_every_n_steps: int, burn_in_steps: int) -> bool:
    return bool(save_trajectory) and (i % save_every_n_steps) == 0 and i >= burn_in_steps


def _single_chain_ptr(y: torch.Tensor) -> torch.Tensor:
    return torch.tensor([0, y.shape[0]], dtype=torch.int32, device=y.device)


def baoab(y: torch.Tensor, score_fn: Callable, steps: int, v_init: Union[str, torch.Tensor] = "zero",
          save_trajectory=False, save_every_n_steps=1, burn_in_steps=0, verbose=False, cpu_offload=False,
          delta: float = 1.0, friction: float = 1.0, M: float = 1.0, inverse_temperature: float = 1.0,
          score_fn_clip: Optional[float] = None, noise: Optional[torch.Tensor] = None, **_):
    """BAOAB with an arbitrary score callable.  ``noise`` ([steps, N, 3]; row 0 feeds v_init) is for parity tests."""
    if isinstance(v_init, str) and v_init not in ("gaussian", "zero"):
        raise RuntimeError(f"{v_init} not in (gaussian, zero)")
    y = y.contiguous().clone()
    sid = next(_call_counter)
    u = pow(M, -1)
    v = initialize_velocity(v_init, y, u, None if noise is None else noise[0].contiguous(), sid)
    prm = _walk_params(delta, friction, M, inverse_temperature, score_fn_clip)
    chain_ptr = _single_chain_ptr(y)
    ybar, p = torch.empty_like(y), torch.empty_like(y)
    y_traj = [] if save_trajectory else None
    score_traj = []
    keep = (lambda t: t.detach().cpu()) if cpu_offload else (lambda t: t.detach().clone())
    for i in range(steps):
        prm.first, prm.last = int(i == 0), int(i == steps - 1)
        prm.step = (sid << 32) + i + 1
        sc = score_fn(y).to(dtype=y.dtype).contiguous()
        if i == 0 or _saves(i, save_trajectory, save_every_n_steps, burn_in_steps):
            if y_traj is not None and _saves(i, save_trajectory, save_every_n_steps, burn_in_steps):
                y_traj.append(keep(y))
            score_traj.append(keep(sc))
        nz = None if noise is None or i + 1 >= noise.shape[0] else noise[i + 1].contiguous()
        ops.walk_step(y, v, ybar, p, None, chain_ptr, prm, nz, None, None, score_in=sc)
    y_traj = torch.stack(y_traj) if y_traj is not None else None
    score_traj = torch.stack(score_traj)
    return y, v, y_traj, score_traj


def aboba(y: torch.Tensor, score_fn: Callable, steps: int, v_init: Union[str, torch.Tensor] = "zero",
          save_trajectory=False, save_every_n_steps=1, burn_in_steps=0, verbose=False, cpu_offload=False,
          delta: float = 1.0, friction: float = 1.0, M: float = 1.0, inverse_temperature: float = 1.0,
          score_fn_clip: Optional[float] = None, noise: Optional[torch.Tensor] = None, **_):
    """ABOBA (score at the half step).  Unlike the reference it does not crash with save_trajectory=False."""
    if isinstance(v_init, str) and v_init not in ("gaussian", "zero"):
        raise RuntimeError(f"{v_init} not in (gaussian, zero)")
    y = y.contiguous().clone()
    sid = next(_call_counter)
    u = pow(M, -1)
    v = initialize_velocity(v_init, y, u, None if noise is None else noise[0].contiguous(), sid)
    prm = _walk_params(delta, friction, M, inverse_temperature, score_fn_clip)
    keep = (lambda t: t.detach().cpu()) if cpu_offload else (lambda t: t.detach().clone())
    y_traj = [] if save_trajectory else None
    if y_traj is not None and 0 >= burn_in_steps:
        y_traj.append(keep(y))
    score_traj = []
    for i in range(1, steps):
        ops.aboba_drift(y, v, delta / 2)
        sc = score_fn(y).to(dtype=y.dtype).contiguous()
        prm.step = (sid << 32) + i
        nz = None if noise is None or i >= noise.shape[0] else noise[i].contiguous()
        ops.aboba_kick(y, v, sc, prm, nz)
        if _saves(i, save_trajectory, save_every_n_steps, burn_in_steps):
            y_traj.append(keep(y))
            score_traj.append(keep(sc))
    y_traj = torch.stack(y_traj) if y_traj is not None else None
    score_traj = torch.stack(score_traj) if score_traj else None
    return y, v, y_traj, score_traj


def fused_baoab(model, topo: "engine.Topology", y: torch.Tensor, sigma: float, steps: int,
                v_init: Union[str, torch.Tensor] = "zero", save_trajectory=False, save_every_n_steps=1, burn_in_steps=0,
                verbose=False, cpu_offload=False, delta: float = 1.0, friction: float = 1.0, M: float = 1.0,
                inverse_temperature: float = 1.0, score_fn_clip: Optional[float] = None,
                noise: Optional[torch.Tensor] = None, **_):
    """Walk-jump hot loop: per step [K1 radius CSR, K2 edge features, 6x(radial MLP, conv, tail), head, K6 step].

    Returns dict(y, v, xhat, y_traj, xhat_traj, score_traj).  xhat / xhat_traj are the jumps at the final and at
    every saved y -- by-products of the score evaluation, not a second pass."""
    if isinstance(v_init, str) and v_init not in ("gaussian", "zero"):
        raise RuntimeError(f"{v_init} not in (gaussian, zero)")
    ctx = model.sigma_context(sigma)
    plan = model.arch_module.plan(ctx.c_noise, y.device)
    mu, rb_step = plan.radial_grid(ctx.r_cut)
    y = y.contiguous().clone()
    sid = next(_call_counter)
    u = pow(M, -1)
    v = initialize_velocity(v_init, y, u, None if noise is None else noise[0].contiguous(), sid)
    prm = _walk_params(delta, friction, M, inverse_temperature, score_fn_clip)
    prm.c_in, prm.c_skip, prm.c_out, prm.sigma2 = ctx.c_in, ctx.c_skip, ctx.c_out, ctx.sigma2
    prm.center = int(model.mean_center)
    N = y.shape[0]
    saved = [i for i in range(steps) if _saves(i, save_trajectory, save_every_n_steps, burn_in_steps)]
    slot = {i: k for k, i in enumerate(saved)}
    score_first_extra = bool(save_trajectory) and 0 not in slot  # reference keeps score(y_init) regardless of burn-in
    T = len(saved)
    f32 = dict(dtype=torch.float32, device=y.device)
    y_traj = torch.empty(T, N, 3, **f32) if save_trajectory else None
    xhat_traj = torch.empty(T, N, 3, **f32) if save_trajectory else None
    score_traj = torch.empty(T + int(score_first_extra), N, 3, **f32) if save_trajectory else torch.empty(1, N, 3, **f32)
    off = int(score_first_extra)
    xhat, score, g = torch.empty_like(y), torch.empty_like(y), torch.empty_like(y)
    ybar, p = ops.center_scale(y, topo.chain_ptr, ctx.c_in, center=model.mean_center)
    for i in range(steps):
        topo.build_csr(ybar, ctx.r_cut)
        engine.e3conv_forward(plan, topo, p, ctx.r_cut, g, mu, rb_step)
        prm.first, prm.last = int(i == 0), int(i == steps - 1)
        prm.step = (sid << 32) + i + 1
        k = slot.get(i)
        ty = y_traj[k] if k is not None else None
        tx = xhat_traj[k] if k is not None else None
        ts = score_traj[k + off] if k is not None else (score_traj[0] if i == 0 else None)
        nz = None if noise is None or i + 1 >= noise.shape[0] else noise[i + 1].contiguous()
        ops.walk_step(y, v, ybar, p, g, topo.chain_ptr, prm, nz, xhat, score, ty, tx, ts)
    if cpu_offload:
        y_traj = y_traj.cpu() if y_traj is not None else None
        xhat_traj = xhat_traj.cpu() if xhat_traj is not None else None
        score_traj = score_traj.cpu()
    return {"y": y, "v": v, "xhat": xhat, "y_traj": y_traj, "xhat_traj": xhat_traj, "score_traj": score_traj}
