"""`average_squared_distance` normalisation constant (mirror of /root/reference/src/jamun/utils/average_squared_distance.py:
139-177): the mean over graphs of the mean squared pair distance below the cutoff -- one reduction kernel per batch."""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from .. import ops


def compute_average_squared_distance(x, cutoff: Optional[float] = None, chain_ptr: Optional[torch.Tensor] = None) -> float:
    """Mean squared distance between the points of one graph (or, with chain_ptr, the mean over graphs of that quantity)."""
    x = torch.as_tensor(x, dtype=torch.float32)
    if not x.is_cuda:
        raise RuntimeError("jamun_b200 compute_average_squared_distance runs on CUDA tensors only (no CPU fallback)")
    if chain_ptr is None:
        chain_ptr = torch.tensor([0, x.shape[0]], dtype=torch.int32, device=x.device)
    sums = ops.avg_sq_dist(x.contiguous(), chain_ptr, -1.0 if cutoff is None else float(cutoff)).cpu()
    per_graph = sums[:, 0] / sums[:, 1].clamp_min(1.0)
    return float(per_graph.mean())


def compute_average_squared_distance_from_data(batches: Iterable, cutoff: float, num_estimation_graphs: int = 5000) -> float:
    """The reference's estimation loop over a dataloader of Batches."""
    total, n = 0.0, 0
    for batch in batches:
        ptr = batch.ptr.to(batch.pos.device, torch.int32)
        sums = ops.avg_sq_dist(batch.pos.float().contiguous(), ptr, float(cutoff)).cpu()
        total += float((sums[:, 0] / sums[:, 1].clamp_min(1.0)).sum())
        n += sums.shape[0]
        if n >= num_estimation_graphs:
            break
    return total / max(n, 1)
