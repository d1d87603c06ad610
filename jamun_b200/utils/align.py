"""Batched Kabsch alignment (mirror of /root/reference/src/jamun/utils/align.py:9-56,123-126).

Training-side only (runs under no_grad before the denoiser call).  One kernel, one warp per chain: centroids, the 3x3
covariance, an in-register one-sided Jacobi SVD in fp64 and the rigid transform (``jamun_kabsch_align``)."""
from __future__ import annotations

import torch


def _chain_ptr(batch: torch.Tensor, num_graphs: int) -> torch.Tensor:
    counts = torch.bincount(batch.to("cpu", torch.long), minlength=num_graphs)
    ptr = torch.zeros(num_graphs + 1, dtype=torch.int32)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr


def kabsch_algorithm(y: torch.Tensor, x: torch.Tensor, batch: torch.Tensor, num_graphs: int, chain_ptr: torch.Tensor = None
                     ) -> torch.Tensor:
    """y aligned onto x per graph: R y + t with R = V diag(1, 1, det) U^T from the SVD of sum y_c x_c^T."""
    from .. import autograd_ops  # noqa: F401

    if not y.is_cuda:
        raise RuntimeError("jamun_b200 kabsch_algorithm runs on CUDA tensors only (no CPU fallback)")
    if chain_ptr is None:
        chain_ptr = _chain_ptr(batch, num_graphs).to(y.device)
    return torch.ops.jamun_b200.kabsch_align(y.float().contiguous(), x.float().contiguous(), chain_ptr)


def align_A_to_B_batched(A, B):
    topo = A["_topology"] if "_topology" in A else None
    ptr = topo.chain_ptr if topo is not None and topo.device == A.pos.device else None
    A.pos = kabsch_algorithm(A.pos, B.pos, A.batch.to(A.pos.device), A.num_graphs, chain_ptr=ptr)
    return A
