"""TEST INFRASTRUCTURE (round 1's library-path training forward, kept as the torch/autograd reference of the custom ops in
jamun_b200/autograd_ops.py -- nothing under jamun_b200/ imports it).

Differentiable evaluation of the denoiser for training.

Status: LIBRARY PATH.  The sampling path runs on this repo's CUDA kernels; their backward counterparts (dM = A^T dOut and
dA = dOut M^T on tcgen05, the transposed aggregate builder) are not written yet.  Until they are, ``Denoiser.training_step``
evaluates the *same restructured math* (aggregate-then-transform over the packed operands of ``Conv.pack`` /
``ConvBlock.pack`` / ``Linear.packed``, SoA irreps layout -- DESIGN.md 3) with torch operators on the GPU, so that autograd
provides the gradients.  The neighbour list still comes from ``jamun_radius_csr`` (integers, no gradient).  The per-receiver
aggregate is a batched GEMM over degree-padded edge blocks, never a per-edge [E, 65, U] tensor.

CUDA tensors only: there is no CPU fallback (the forward agrees with the kernel path to fp32 round-off and is tested against it
and against the oracle's autograd gradients in tests/test_gpu_train.py).
Mirrors /root/reference/src/jamun/model/denoiser.py:168-217 (xhat_normalized / xhat) and arch/e3conv.py:87-138.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from jamun_b200 import engine, ops

S, V, HID, SO = ops.S, ops.V, ops.HID, 152


def _noise_mlp(mod, c_noise: float, sigmoid: bool) -> torch.Tensor:
    w1, b1, w2, b2 = mod.mlp_operands()
    out = F.selu(w1 * c_noise + b1) @ w2.T + b2
    return torch.sigmoid(out) if sigmoid else out


def _expand(w: torch.Tensor) -> torch.Tensor:  # per-irrep [152] -> SoA [216]
    return torch.cat([w[:S], w[S:], w[S:], w[S:]])


def _padded_edge_index(rowptr: torch.Tensor, E: int):
    """[N, D] edge ids of every receiver's in-edges, padded with the dummy id E."""
    deg = (rowptr[1:] - rowptr[:-1]).long()
    D = int(deg.max().item()) if deg.numel() else 0
    ar = torch.arange(max(D, 1), device=rowptr.device)
    eid = rowptr[:-1].long()[:, None] + ar[None, :]
    return torch.where(ar[None, :] < deg[:, None], eid, torch.full_like(eid, E)), deg


def _conv(x, s_in, v_in, src, eid_pad, deg, h, rhat, pk):
    """Aggregate-then-transform conv on SoA features: x [N, s_in + 3 v_in] -> [N, 248]."""
    E, N = src.shape[0], x.shape[0]
    hp = torch.cat([h, torch.ones(E, 1, dtype=h.dtype, device=h.device)], dim=1)  # [E, 65]
    xs = x[src, :s_in]
    f0 = [xs]
    f1 = [[xs * rhat[:, c:c + 1]] for c in range(3)]
    if v_in:
        xv = x[src, s_in:].reshape(E, 3, v_in)
        f0.append((xv * rhat[:, :, None]).sum(1))
        cross = torch.cross(xv, rhat[:, :, None].expand(E, 3, v_in), dim=1)
        for c in range(3):
            f1[c] += [xv[:, c] / math.sqrt(3.0), cross[:, c] / math.sqrt(2.0)]
    zero_row = lambda t: torch.cat([t, t.new_zeros(1, t.shape[1])], dim=0)  # noqa: E731  (dummy edge E)
    hp_pad = zero_row(hp)[eid_pad]                                           # [N, D, 65]
    inv = (1.0 / deg.clamp_min(1).to(x.dtype))[:, None]

    def contract(f, m, alpha):
        fp = zero_row(f)[eid_pad]                                            # [N, D, U]
        A = torch.bmm(hp_pad.transpose(1, 2), fp)                            # [N, 65, U]  (per-receiver aggregate)
        return (A.reshape(N, -1) @ m.reshape(-1, m.shape[2])) * alpha * inv

    outs = [contract(torch.cat(f0, dim=1), pk["m0"], pk["alpha0"])]
    for c in range(3):
        outs.append(contract(torch.cat(f1[c], dim=1), pk["m1"], pk["alpha1"]))
    return torch.cat(outs, dim=1)


def _block_tail(conv_out, x_in, s_in, v_in, x_res, pk, skip_w, s_next):
    N = conv_out.shape[0]
    gs = pk["c_act"] * F.leaky_relu(conv_out[:, :S], 0.01)
    gate = pk["c_gate"] * torch.sigmoid(conv_out[:, S:SO])
    gv = conv_out[:, SO:].reshape(N, 3, V) * gate[:, None, :]
    ys = gs @ pk["wself_s"] + x_in[:, :s_in] @ pk["wskip_s"]
    yv = gv @ pk["wself_v"]
    if v_in:
        yv = yv + x_in[:, s_in:].reshape(N, 3, v_in) @ pk["wskip_v"]
    y = torch.cat([ys, yv.reshape(N, 3 * V)], dim=1)
    if skip_w is not None:
        w = _expand(skip_w)
        y = x_res * w + y * (1 - w)
    return y, (y * _expand(s_next) if s_next is not None else y)


def network_output(arch, topo: engine.Topology, p: torch.Tensor, c_noise: float, r_cut: float) -> torch.Tensor:
    """g = E3Conv(p) [N, 3] on topo's current CSR, differentiable w.r.t. arch's parameters."""
    if not p.is_cuda:
        raise RuntimeError("jamun_b200.train runs on CUDA tensors only (no CPU fallback)")
    E = int(topo.rowptr[-1].item())
    src, dst = topo.col[:E].long(), topo.edst[:E].long()
    ebond = topo.ebond[:E].long()
    with torch.no_grad():
        vec = p[src] - p[dst]
        d = vec.norm(dim=1)
        rhat = vec / d.clamp_min(1e-12)[:, None]
        values = torch.linspace(0.0, float(r_cut), ops.NBASIS + 2, dtype=p.dtype, device=p.device)
        rb = (-(((d[:, None] - values[1:-1]) / (values[1] - values[0])) ** 2)).exp() / 1.12
        eid_pad, deg = _padded_edge_index(topo.rowptr, E)
    emb = arch.embed_bondedness.weight
    tables = arch.atom_embedder.tables()
    idx = list(topo.idx)
    if not arch.atom_embedder.use_residue_sequence_index:
        idx[3] = None
    cols = [t[(i.long() if i is not None else torch.zeros(topo.N, dtype=torch.long, device=p.device))] for t, i in zip(tables, idx)]
    x0 = torch.cat(cols, dim=1)
    s_init = _noise_mlp(arch.initial_noise_scaling, c_noise, False)
    x_in, x_res = x0 * s_init, None
    blocks = [arch.initial_projector, *arch.layers]
    nb = len(blocks)
    for l, blk in enumerate(blocks):
        pk = blk.pack(emb)
        z = rb @ pk["w0r"] + pk["b0eff"][ebond]
        h = z * torch.sigmoid(z)
        conv_out = _conv(x_in, pk["s_in"], pk["v_in"], src, eid_pad, deg, h, rhat, pk)
        skip_w = _noise_mlp(arch.skip_connections[l - 1].weights, c_noise, True) if l > 0 else None
        s_next = _noise_mlp(arch.noise_scalings[l], c_noise, False) if l < nb - 1 else None
        x_new, x_scaled = _block_tail(conv_out, x_in, pk["s_in"], pk["v_in"], x_res, pk, skip_w, s_next)
        x_in, x_res = x_scaled, x_new
    hb, lin2 = arch.output_head[0], arch.output_head[1]
    w1s, w1v = hb.lin.packed(0), hb.lin.packed(1)
    w2 = lin2.packed(1).reshape(-1) * arch.output_gain
    N = x_res.shape[0]
    gate = hb.gate.c_gate * torch.sigmoid(x_res[:, :S] @ w1s[:, S:])
    hv = x_res[:, S:].reshape(N, 3, V) @ w1v
    return ((hv * gate[:, None, :]) * w2).sum(-1)


def xhat_positions(denoiser, y: torch.Tensor, topo: engine.Topology, sigma) -> torch.Tensor:
    """xhat [N, 3] with an autograd graph to the denoiser's parameters (denoiser.py:168-217)."""
    ctx = denoiser.sigma_context(sigma)
    with torch.no_grad():
        if denoiser.mean_center:
            ybar, p = ops.center_scale(y.contiguous(), topo.chain_ptr, ctx.c_in)
        else:
            ybar, p = y, y * ctx.c_in
        topo.build_csr(ybar, ctx.r_cut)
    g = network_output(denoiser.arch_module, topo, p, ctx.c_noise, ctx.r_cut)
    xh = ctx.c_skip * ybar + ctx.c_out * g
    if denoiser.mean_center:
        cnt = (topo.chain_ptr_long[1:] - topo.chain_ptr_long[:-1]).clamp_min(1).to(xh.dtype)
        mean = torch.zeros(topo.G, 3, dtype=xh.dtype, device=xh.device).index_add_(0, topo.batch_long, xh) / cnt[:, None]
        xh = xh - mean[topo.batch_long]
    return xh
