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
import os
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


class _WalkWorkspace:
    """Persistent device buffers of the fused walk (so CUDA graphs captured once can be replayed across calls)."""

    def __init__(self, N: int, T: int, score_rows: int, save: bool, device):
        f32 = dict(dtype=torch.float32, device=device)
        self.y, self.v, self.ybar, self.p, self.g, self.xhat, self.score = (torch.empty(N, 3, **f32) for _ in range(7))
        self.y_traj = torch.empty(T, N, 3, **f32) if save else None
        self.xhat_traj = torch.empty(T, N, 3, **f32) if save else None
        self.score_traj = torch.empty(score_rows, N, 3, **f32)
        self.state = torch.zeros(2, dtype=torch.int64, device=device)  # [philox step, trajectory slot]
        self.graphs = {}
        self.graph_launches = {}  # kernels per replay of each captured graph
        self.graph_refs = {}      # tensors whose raw pointers a captured graph holds through the C ABI (kept alive with it)


def fused_baoab(model, topo: "engine.Topology", y: torch.Tensor, sigma: float, steps: int,
                v_init: Union[str, torch.Tensor] = "zero", save_trajectory=False, save_every_n_steps=1, burn_in_steps=0,
                verbose=False, cpu_offload=False, delta: float = 1.0, friction: float = 1.0, M: float = 1.0,
                inverse_temperature: float = 1.0, score_fn_clip: Optional[float] = None,
                noise: Optional[torch.Tensor] = None, use_cuda_graph: Optional[bool] = None, **_):
    """Walk-jump hot loop: per step [K1 radius CSR, K2 edge features, 6x(radial MLP, transform, aggregate, GEMM, tail),
    head, K6 step].  Steady-state steps are replayed from a CUDA graph (step counter / trajectory slot in device memory).

    Returns dict(y, v, xhat, y_traj, xhat_traj, score_traj).  xhat / xhat_traj are the jumps at the final and at
    every saved y -- by-products of the score evaluation, not a second pass."""
    if isinstance(v_init, str) and v_init not in ("gaussian", "zero"):
        raise RuntimeError(f"{v_init} not in (gaussian, zero)")
    ctx = model.sigma_context(sigma)
    plan = model.arch_module.plan(ctx.c_noise, y.device)
    mu, rb_step = plan.radial_grid(ctx.r_cut)
    sid = next(_call_counter)
    u = pow(M, -1)
    N = y.shape[0]
    saved = [i for i in range(steps) if _saves(i, save_trajectory, save_every_n_steps, burn_in_steps)]
    slot = {i: k for k, i in enumerate(saved)}
    score_first_extra = bool(save_trajectory) and 0 not in slot  # reference keeps score(y_init) regardless of burn-in
    T = len(saved)
    off = int(score_first_extra)
    score_rows = T + off if save_trajectory else 1
    key = (T, score_rows, bool(save_trajectory))
    cache = topo.__dict__.setdefault("_walk_ws", {})
    ws = cache.get(key)
    if ws is None:
        if len(cache) > 4:
            cache.clear()
        ws = cache[key] = _WalkWorkspace(N, T, score_rows, bool(save_trajectory), y.device)
    ws.y.copy_(y)
    ws.v.copy_(initialize_velocity(v_init, ws.y, u, None if noise is None else noise[0].contiguous(), sid))
    prm = _walk_params(delta, friction, M, inverse_temperature, score_fn_clip)
    prm.c_in, prm.c_skip, prm.c_out, prm.sigma2 = ctx.c_in, ctx.c_skip, ctx.c_out, ctx.sigma2
    prm.center = int(model.mean_center)
    score_view = ws.score_traj[off:] if save_trajectory else None

    def denoise():
        topo.build_csr(ws.ybar, ctx.r_cut)
        engine.e3conv_forward(plan, topo, ws.p, ctx.r_cut, ws.g, mu, rb_step)

    def eager_step(i):
        denoise()
        prm.first, prm.last = int(i == 0), int(i == steps - 1)
        prm.step = (sid << 32) + i + 1
        k = slot.get(i)
        ty = ws.y_traj[k] if k is not None else None
        tx = ws.xhat_traj[k] if k is not None else None
        ts = ws.score_traj[k + off] if k is not None else (ws.score_traj[0] if i == 0 else None)
        nz = None if noise is None or i + 1 >= noise.shape[0] else noise[i + 1].contiguous()
        ops.walk_step(ws.y, ws.v, ws.ybar, ws.p, ws.g, topo.chain_ptr, prm, nz, ws.xhat, ws.score, ty, tx, ts)

    ops.center_scale(ws.y, topo.chain_ptr, ctx.c_in, ws.ybar, ws.p, center=model.mean_center)
    if use_cuda_graph is None:
        use_cuda_graph = noise is None and steps >= 4 and os.environ.get("JAMUN_B200_GRAPH", "1") != "0"
    if not use_cuda_graph:
        for i in range(steps):
            eager_step(i)
    else:
        eager_step(0)  # also warms up every lazily allocated workspace before capture
        gkey = (plan.serial, float(delta), float(friction), float(M), float(inverse_temperature), score_fn_clip,
                float(sigma), int(model.mean_center), int(prm.seed))

        def graph_for(save: bool):
            g = ws.graphs.get((gkey, save))
            if g is None:
                prm.first, prm.last = 0, 0
                g = torch.cuda.CUDAGraph()
                n0 = ops.LAUNCHES
                with torch.cuda.graph(g):
                    denoise()
                    ops.walk_step(ws.y, ws.v, ws.ybar, ws.p, ws.g, topo.chain_ptr, prm, None, ws.xhat, ws.score,
                                  ws.y_traj if save else None, ws.xhat_traj if save else None, score_view if save else None,
                                  dev_state=ws.state)
                    ops.walk_advance(ws.state, 1 if save else 0)
                ws.graphs[(gkey, save)] = g
                ws.graph_refs[(gkey, save)] = (plan, mu)
                ws.graph_launches[(gkey, save)] = ops.LAUNCHES - n0  # kernels recorded in the graph = launched per replay
                ops.LAUNCHES = n0                                     # capture itself launches nothing
            return g

        first_slot = 1 if 0 in slot else 0
        ws.state.copy_(torch.tensor([(sid << 32) + 2, first_slot], dtype=torch.int64))
        if steps > 2:
            # capture (does not execute) before the replay loop so the device state is untouched by capture
            needed = {i in slot for i in range(1, steps - 1)}
            graphs = {s: graph_for(s) for s in needed}
            for i in range(1, steps - 1):
                graphs[i in slot].replay()
                ops._count(ws.graph_launches[(gkey, i in slot)])
        if steps > 1:
            eager_step(steps - 1)
    if plan.gemm_kind == "f16" and topo.overflowed():  # 4-byte read-back of the fp16-split GEMMs' range flag
        engine.fall_back_to_tf32("the walk")
        return fused_baoab(model, topo, y, sigma, steps, v_init=v_init, save_trajectory=save_trajectory,
                           save_every_n_steps=save_every_n_steps, burn_in_steps=burn_in_steps, verbose=verbose, cpu_offload=cpu_offload,
                           delta=delta, friction=friction, M=M, inverse_temperature=inverse_temperature, score_fn_clip=score_fn_clip,
                           noise=noise, use_cuda_graph=use_cuda_graph)
    out_y, out_v, out_x = ws.y.clone(), ws.v.clone(), ws.xhat.clone()
    y_traj = ws.y_traj.clone() if ws.y_traj is not None else None
    xhat_traj = ws.xhat_traj.clone() if ws.xhat_traj is not None else None
    score_traj = ws.score_traj.clone()
    if cpu_offload:
        y_traj = y_traj.cpu() if y_traj is not None else None
        xhat_traj = xhat_traj.cpu() if xhat_traj is not None else None
        score_traj = score_traj.cpu()
    return {"y": out_y, "v": out_v, "xhat": out_x, "y_traj": y_traj, "xhat_traj": xhat_traj, "score_traj": score_traj}
