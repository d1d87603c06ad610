"""Torch (CPU) emulation of the *math* each CUDA kernel implements, operating on the product's packed operands.

Test infrastructure only: it lets the CPU suite check the algebraic restructuring (aggregate-then-transform) and the
weight re-layout (Conv.pack / Linear.packed / E3ConvPlan) against the oracle without a GPU; the GPU suite then checks
the kernels against the oracle directly.
"""
from __future__ import annotations

import math

import torch

S, V, HID, SO = 120, 32, 216, 152


def to_soa(x, s, v):
    n = x.shape[0]
    return torch.cat([x[:, :s], x[:, s:].reshape(n, v, 3).permute(0, 2, 1).reshape(n, 3 * v)], dim=1)


def from_soa(x, s, v):
    n = x.shape[0]
    return torch.cat([x[:, :s], x[:, s:].reshape(n, 3, v).permute(0, 2, 1).reshape(n, 3 * v)], dim=1)


def edge_geom(p, src, dst, r_cut, n_basis=32):
    vec = p[src] - p[dst]
    d = vec.norm(dim=1)
    rhat = vec / d.clamp_min(1e-12)[:, None]
    values = torch.linspace(0.0, float(r_cut), n_basis + 2, dtype=p.dtype)
    step = values[1] - values[0]
    rb = (-(((d[:, None] - values[1:-1]) / step) ** 2)).exp() / 1.12
    return rhat, rb


def radial_hidden(rb, ebond, w0r, b0eff):
    z = rb @ w0r + b0eff[ebond.long()]
    return z * torch.sigmoid(z)


def conv(x, s_in, v_in, src, dst, h, rhat, m0, m1, alpha0, alpha1, N):
    """x: SoA [N, s_in+3 v_in] -> [N, 248] SoA."""
    E = src.shape[0]
    hp = torch.cat([h, torch.ones(E, 1, dtype=h.dtype)], dim=1)  # [E,65]
    xs = x[src, :s_in]
    f0 = [xs]
    f1 = [[xs * rhat[:, c:c + 1]] for c in range(3)]
    if v_in:
        xv = x[src, s_in:].reshape(E, 3, v_in)
        f0.append((xv * rhat[:, :, None]).sum(1))
        cross = torch.cross(xv, rhat[:, :, None].expand(E, 3, v_in), dim=1)
        for c in range(3):
            f1[c] += [xv[:, c] / math.sqrt(3.0), cross[:, c] / math.sqrt(2.0)]
    f0 = torch.cat(f0, dim=1)
    f1 = [torch.cat(t, dim=1) for t in f1]
    deg = torch.zeros(N, dtype=x.dtype).index_add_(0, dst, torch.ones(E, dtype=x.dtype)).clamp_min(1.0)
    A0 = torch.zeros(N, 65, f0.shape[1], dtype=x.dtype).index_add_(0, dst, hp[:, :, None] * f0[:, None, :])
    out0 = torch.einsum("nku,kuw->nw", A0, m0) * alpha0 / deg[:, None]
    outs = [out0]
    for c in range(3):
        A1 = torch.zeros(N, 65, f1[c].shape[1], dtype=x.dtype).index_add_(0, dst, hp[:, :, None] * f1[c][:, None, :])
        outs.append(torch.einsum("nku,kuw->nw", A1, m1) * alpha1 / deg[:, None])
    return torch.cat(outs, dim=1)


def _expand(w):  # per-irrep [152] -> SoA [216]
    return torch.cat([w[:S], w[S:], w[S:], w[S:]])


def block_tail(conv_out, x_in, s_in, v_in, x_res, b, skip_w, s_next):
    N = conv_out.shape[0]
    gs = b["c_act"] * torch.nn.functional.leaky_relu(conv_out[:, :S], 0.01)
    gate = b["c_gate"] * torch.sigmoid(conv_out[:, S:SO])
    gv = conv_out[:, SO:].reshape(N, 3, V) * gate[:, None, :]
    ys = gs @ b["wself_s"] + x_in[:, :s_in] @ b["wskip_s"]
    yv = gv @ b["wself_v"]
    if v_in:
        yv = yv + x_in[:, s_in:].reshape(N, 3, v_in) @ b["wskip_v"]
    y = torch.cat([ys, yv.reshape(N, 3 * V)], dim=1)
    if skip_w is not None:
        w = _expand(skip_w)
        y = x_res * w + y * (1 - w)
    return y, (y * _expand(s_next) if s_next is not None else y)


def head(x, w1s, w1v, w2, c_gate):
    N = x.shape[0]
    gate = c_gate * torch.sigmoid(x[:, :S] @ w1s[:, S:])
    hv = x[:, S:].reshape(N, 3, V) @ w1v
    return ((hv * gate[:, None, :]) * w2).sum(-1)


def noise_mlp(w1, b1, w2, b2, c, sigmoid):
    out = torch.nn.functional.selu(w1 * c + b1) @ w2.T + b2
    return torch.sigmoid(out) if sigmoid else out
