"""Data-parallel training: bucketed gradient all-reduce over NVLink, overlapped with the backward pass.

The reference trains with Lightning's DDP strategy (``/root/reference/src/jamun/hydra_config/trainer/default.yaml:1``,
``scripts/slurm/train.sh:29-35``): one process per GPU, identical replicas, gradients averaged over ranks every step.  Here
one :class:`GradientReducer` per model does the same with ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests):

* gradients live in flat per-bucket buffers (``p.grad`` are views), so a bucket is reduced in place with no packing copy;
* buckets follow the order in which gradients become ready in backward (output head first, initial block last).  Each of
  the five 1.86 M-element ``radial_nn.3.weight`` tensors -- 88 % of the 42.2 MB of gradients -- is its own bucket, so its
  all-reduce starts as soon as that layer's ``jamun_stage_atb`` weight-gradient kernel has run and overlaps the backward
  kernels of the layers below it;
* the collective is enqueued on NCCL's stream (``async_op=True``) from a post-accumulate-grad hook; ``finish()`` waits for
  all buckets and leaves the averaged gradients in ``p.grad`` for the optimiser.

Chains (graphs) are independent, so this all-reduce is the only exchange of a training step.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class GradientReducer:
    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 4 << 20, process_group=None):
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        # NCCL averages inside the collective; gloo (CPU tests) sums and the mean is taken afterwards
        self._avg = self.world > 1 and dist.get_backend(process_group) == "nccl"
        params = [p for p in module.parameters() if p.requires_grad]
        # reverse registration order ~ order in which backward produces the gradients
        order = list(reversed(params))
        self.buckets: List[Dict] = []
        cur, cur_bytes = [], 0
        for p in order:
            nbytes = p.numel() * p.element_size()
            if nbytes >= bucket_bytes:  # a large tensor is its own bucket: reduce it the moment it is ready
                if cur:
                    self._close(cur)
                    cur, cur_bytes = [], 0
                self._close([p])
                continue
            cur.append(p)
            cur_bytes += nbytes
            if cur_bytes >= bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._close(cur)
        self._bucket_of = {}
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._bucket_of[p] = bi
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.reset()

    def _close(self, params: List[torch.nn.Parameter]) -> None:
        p0 = params[0]
        flat = torch.zeros(sum(p.numel() for p in params), dtype=p0.dtype, device=p0.device)
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p)  # gradients accumulate straight into the bucket
            off += p.numel()
        self.buckets.append({"params": list(params), "flat": flat, "pending": 0, "work": None})

    # ------------------------------------------------------------------ per step
    def reset(self) -> None:
        """Zero the gradient buckets (instead of optimizer.zero_grad(set_to_none=True), which would detach the views)."""
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            b["work"] = None
            off = 0
            for p in b["params"]:  # re-attach views an optimiser may have replaced
                view = b["flat"][off:off + p.numel()].view_as(p)
                if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                    p.grad = view
                off += p.numel()

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        b = self.buckets[self._bucket_of[p]]
        if p.grad.data_ptr() < b["flat"].data_ptr() or p.grad.data_ptr() >= b["flat"].data_ptr() + b["flat"].numel() * 4:
            # autograd replaced the view (first accumulation into a None grad): copy into the bucket and re-attach
            off = 0
            for q in b["params"]:
                if q is p:
                    view = b["flat"][off:off + p.numel()].view_as(p)
                    view.copy_(p.grad)
                    p.grad = view
                off += q.numel()
        b["pending"] -= 1
        if b["pending"] == 0 and self.world > 1:
            b["work"] = self._all_reduce(b["flat"])

    def _all_reduce(self, flat: torch.Tensor):
        return dist.all_reduce(flat, op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> None:
        """Wait for every bucket and turn the sums into means.  Parameters that received no gradient this step (unused
        branches) still take part so that all ranks issue the same collectives."""
        for b in self.buckets:
            if self.world > 1 and b["work"] is None:
                b["work"] = self._all_reduce(b["flat"])
        for b in self.buckets:
            if b["work"] is not None:
                b["work"].wait()
                if not self._avg:
                    b["flat"].div_(self.world)

    @property
    def gradient_bytes(self) -> int:
        return sum(b["flat"].numel() * b["flat"].element_size() for b in self.buckets)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()


def broadcast_parameters(module: torch.nn.Module, src: int = 0, process_group=None) -> None:
    """Replicas start identical (DDP's constructor broadcast)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=process_group)
