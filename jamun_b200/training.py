"""Training-time evaluation of the denoiser: forward and backward on this library's kernels.

``network_output`` composes the torch.library operators of :mod:`jamun_b200.autograd_ops` -- each a sequence of hand-written
CUDA kernels with a hand-written backward -- over the packed operands of ``Conv.pack`` / ``ConvBlock.pack`` /
``Linear.packed``.  Those ``pack`` methods are pure re-layouts of the parameters (slicing, transposition, the 1/sqrt(fan_in)
path weights, the 2x32x64 bondedness fold); torch's autograd maps the operand gradients back onto the reference's
``state_dict`` tensors through them.  Everything batch-sized -- convolution, gate, Linears, reductions, loss, Kabsch --
runs in kernels of this repository.

Mirrors /root/reference/src/jamun/model/denoiser.py:168-217 (xhat), arch/e3conv.py:87-138 (network), denoiser.py:219-319
(noise_and_denoise / compute_loss / training_step).
"""
from __future__ import annotations

import torch

from . import autograd_ops as A  # noqa: F401  (registers torch.ops.jamun_b200.*)
from . import engine, ops

T = torch.ops.jamun_b200


_grids = {}  # (r_cut, device) -> (centres on the device, spacing)


def _radial_grid(r_cut: float, device):
    """Radial-basis centres for this cut-off (host linspace, as the sampling path's plan.radial_grid).  Cached per cut-off; a new
    cut-off (sigma drawn from a continuous distribution: one per step) is uploaded from pinned memory without blocking -- a
    pageable host-to-device copy would wait for the stream and stop the host from running ahead."""
    key = (r_cut, str(device))
    if key not in _grids:
        if len(_grids) > 64:
            _grids.clear()
        values = torch.linspace(0.0, r_cut, ops.NBASIS + 2, dtype=torch.float32)
        _grids[key] = (values[1:-1].contiguous().pin_memory().to(device, non_blocking=True), float(values[1] - values[0]))
    return _grids[key]


def network_output(arch, topo: engine.Topology, p: torch.Tensor, c_noise: float, r_cut: float) -> torch.Tensor:
    """g = E3Conv(p) [N, 3] on topo's current CSR, differentiable w.r.t. arch's parameters."""
    if not p.is_cuda:
        raise RuntimeError("jamun_b200.training runs on CUDA tensors only (no CPU fallback)")
    with torch.no_grad():  # geometry carries no gradient (positions are inputs)
        mu, step = _radial_grid(float(r_cut), p.device)
        ops.edge_geom(p.contiguous(), topo.rowptr, topo.col, topo.edst, mu, step, topo.rhat, topo.rb)
    csr = (topo.rhat, topo.rowptr, topo.col, topo.edst, topo.src_rowptr, topo.src_eid)
    emb = arch.embed_bondedness.weight
    idx = list(topo.idx)
    if not arch.atom_embedder.use_residue_sequence_index:
        idx[3] = None
    s_init = T.noise_mlp(*arch.initial_noise_scaling.mlp_operands(), c_noise, False)
    x_in = T.atom_embed(*idx, *arch.atom_embedder.tables(), s_init)
    x_res = None
    blocks = [arch.initial_projector, *arch.layers]
    nb = len(blocks)
    for l, blk in enumerate(blocks):
        pk = blk.pack(emb)
        h = T.radial_hidden(topo.rb, topo.ebond, topo.rowptr, pk["w0r"], pk["b0eff"])
        conv_out = T.conv_keep(x_in, h, *csr, pk["m0"], pk["m1"], pk["s_in"], pk["v_in"], pk["alpha0"], pk["alpha1"])[0]
        skip_w = T.noise_mlp(*arch.skip_connections[l - 1].weights.mlp_operands(), c_noise, True) if l > 0 else None
        s_next = T.noise_mlp(*arch.noise_scalings[l].mlp_operands(), c_noise, False) if l < nb - 1 else None
        x_new, x_scaled = T.block_tail(conv_out, x_in, x_res, pk["wself_s"], pk["wself_v"], pk["wskip_s"], pk["wskip_v"], skip_w,
                                       s_next, pk["s_in"], pk["v_in"], pk["c_act"], pk["c_gate"])
        x_in, x_res = x_scaled, x_new
    hb, lin2 = arch.output_head[0], arch.output_head[1]
    w2 = lin2.packed(1).reshape(-1) * arch.output_gain
    return T.head(x_res, hb.lin.packed(0), hb.lin.packed(1), w2, hb.gate.c_gate)


def xhat_positions(denoiser, y: torch.Tensor, topo: engine.Topology, sigma) -> torch.Tensor:
    """xhat [N, 3] with an autograd graph to the denoiser's parameters (denoiser.py:168-217)."""
    ctx = denoiser.sigma_context(sigma)
    with torch.no_grad():
        ybar, p = ops.center_scale(y.contiguous(), topo.chain_ptr, ctx.c_in, center=denoiser.mean_center)
        topo.build_csr(ybar, ctx.r_cut)
    g = network_output(denoiser.arch_module, topo, p, ctx.c_noise, ctx.r_cut)
    return T.combine_xhat(g, ybar, topo.chain_ptr, ctx.c_skip, ctx.c_out, bool(denoiser.mean_center))
