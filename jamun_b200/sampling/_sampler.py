"""Sampler (mirror of /root/reference/src/jamun/sampling/_sampler.py:15-98) without a hard Lightning dependency.

One process per GPU: rank r uses cuda:LOCAL_RANK and its own RNG stream (seed + rank, as cmdline/sample.py:86-88);
chains are independent, so nothing is communicated during the walk.  ``gather_samples`` does the single final
NCCL all-gather of the denoised samples.
"""
from __future__ import annotations

import os
from typing import Any, Iterable, List, Optional, Union

import torch

from .. import utils


class _FabricLite:
    """The members of lightning.Fabric the sampler relies on (launch/setup/device/call/log/rank)."""

    def __init__(self, accelerator="auto", devices="auto", callbacks=None, loggers=None, **_):
        self.callbacks = list(callbacks) if isinstance(callbacks, (list, tuple)) else ([callbacks] if callbacks else [])
        self.loggers = list(loggers) if isinstance(loggers, (list, tuple)) else ([loggers] if loggers else [])
        self.global_rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        if accelerator == "cpu" or not torch.cuda.is_available():
            raise RuntimeError("jamun_b200.Sampler needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", self.local_rank)
        self.logged = {}

    @property
    def is_global_zero(self) -> bool:
        return self.global_rank == 0

    def launch(self):
        torch.cuda.set_device(self.device)

    def setup(self, model):
        return model.to(self.device)

    def call(self, hook_name: str, *args, **kwargs):
        for cb in self.callbacks:
            fn = getattr(cb, hook_name, None)
            if callable(fn):
                fn(*args, **kwargs)

    def log(self, name, value, step=None):
        self.logged[name] = value
        for lg in self.loggers:
            if hasattr(lg, "log_metrics"):
                lg.log_metrics({name: value}, step=step)


class Sampler:
    """A sampler for molecular dynamics simulations."""

    def __init__(self, accelerator: Union[str, Any] = "auto", strategy: Union[str, Any] = "auto",
                 devices: Union[List[int], str, int] = "auto", num_nodes: int = 1, precision: Union[str, int] = "32-true",
                 plugins: Optional[Any] = None, callbacks: Optional[Any] = None, loggers: Optional[Any] = None):
        if str(precision) not in ("32-true", "32", "highest"):
            raise NotImplementedError("the walk-jump kernels compute in fp32 (precision='32-true')")
        self.fabric = _FabricLite(accelerator=accelerator, devices=devices, callbacks=callbacks, loggers=loggers)
        self.global_step = None

    def progbar_wrapper(self, iterable: Iterable, total: int, **kwargs: Any):
        if self.fabric.is_global_zero:
            try:
                from tqdm.auto import tqdm

                return tqdm(iterable, total=total, **kwargs)
            except Exception:  # noqa: BLE001
                return iterable
        return iterable

    def sample(self, model, batch_sampler, num_batches: int, init_graphs, continue_chain: bool = False,
               return_outputs: bool = False):
        """Mirror of _sampler.py:53-98.  Like the reference it returns nothing: every batch's samples go to the callbacks
        (``on_after_sample_batch``).  ``return_outputs=True`` (tests, notebooks) keeps each batch's output dict, moved to
        the host so that long runs do not accumulate trajectories in device memory."""
        self.fabric.launch()
        model = self.fabric.setup(model)
        model.eval()
        init_graphs = init_graphs.to(self.fabric.device)
        model_wrapped = utils.ModelSamplingWrapper(model=model, init_graphs=init_graphs, sigma=batch_sampler.sigma)
        y_init = model_wrapped.sample_initial_noisy_positions()
        v_init = "gaussian"
        self.fabric.call("on_sample_start", sampler=self)
        outputs = [] if return_outputs else None
        with torch.inference_mode():
            for batch_idx in self.progbar_wrapper(range(num_batches), total=num_batches, desc="Sampling", leave=False):
                self.global_step = batch_idx
                self.fabric.call("on_before_sample_batch", sampler=self)  # the reference never fires it
                out = batch_sampler.sample(model=model_wrapped, y_init=y_init, v_init=v_init)
                samples = model_wrapped.unbatch_samples(out)
                if continue_chain:
                    y_init = out["y"].to(model_wrapped.device)
                    v_init = out["v"].to(model_wrapped.device)
                else:
                    y_init = model_wrapped.sample_initial_noisy_positions()
                    v_init = "gaussian"
                self.fabric.call("on_after_sample_batch", sample=samples, sampler=self)
                self.fabric.log("sampler/global_step", batch_idx)
                if outputs is not None:
                    outputs.append({k: (v.detach().cpu() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})
        self.fabric.call("on_sample_end", sampler=self)
        return outputs

    @staticmethod
    def gather_samples(sample: torch.Tensor) -> torch.Tensor:
        """Final exchange of the sharded run: all-gather per-rank samples [N_r, ...] (equal N_r) over NCCL."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return sample
        out = [torch.empty_like(sample) for _ in range(dist.get_world_size())]
        dist.all_gather(out, sample.contiguous())
        return torch.cat(out, dim=0)

    @staticmethod
    def gather_samples_ragged(sample: torch.Tensor) -> torch.Tensor:
        """All-gather of per-rank samples with different atom counts (pads to the largest shard)."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return sample
        world = dist.get_world_size()
        n = torch.tensor([sample.shape[0]], device=sample.device)
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        mx = max(int(t) for t in ns)
        buf = sample.new_zeros((mx,) + tuple(sample.shape[1:]))
        buf[: sample.shape[0]] = sample
        outs = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        return torch.cat([o[: int(k)] for o, k in zip(outs, ns)], dim=0)
