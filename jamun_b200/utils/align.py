"""Batched Kabsch alignment (mirror of /root/reference/src/jamun/utils/align.py:9-56,123-126).

Training-side only (runs under no_grad before the denoiser call).  Round 1: segment reductions via
torch index_add on the device + torch.linalg.svd; a fused per-chain kernel (K7 in DESIGN.md) is the
planned replacement.  Not on the sampling path.
"""
from __future__ import annotations

import torch


def _segment_mean(x, batch, num_graphs):
    s = torch.zeros(num_graphs, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    n = torch.zeros(num_graphs, dtype=x.dtype, device=x.device).index_add_(0, batch, torch.ones_like(x[:, 0]))
    return s / n.clamp_min(1.0)[:, None]


def kabsch_algorithm(y: torch.Tensor, x: torch.Tensor, batch: torch.Tensor, num_graphs: int) -> torch.Tensor:
    x_mu, y_mu = _segment_mean(x, batch, num_graphs), _segment_mean(y, batch, num_graphs)
    x_c, y_c = x - x_mu[batch], y - y_mu[batch]
    H = torch.zeros(num_graphs, 3, 3, dtype=y.dtype, device=y.device).index_add_(
        0, batch, y_c[:, :, None] * x_c[:, None, :])
    U, _, VH = torch.linalg.svd(H)
    R = torch.einsum("Gki,Gjk->Gij", VH, U)
    dets = torch.linalg.det(R)
    signs = torch.eye(3, device=y.device, dtype=y.dtype).repeat(num_graphs, 1, 1)
    signs[:, 2, 2] = dets
    R = torch.einsum("Gki,Gkk,Gjk->Gij", VH, signs, U)
    t = x_mu - torch.einsum("Gij,Gj->Gi", R, y_mu)
    return torch.einsum("Nij,Nj->Ni", R[batch], y) + t[batch]


def align_A_to_B_batched(A, B):
    A.pos = kabsch_algorithm(A.pos, B.pos, A.batch.to(A.pos.device), A.num_graphs)
    return A
