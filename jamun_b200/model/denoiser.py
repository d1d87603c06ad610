"""Denoiser (mirror of /root/reference/src/jamun/model/denoiser.py:13-353) on the sm_100a kernels.

Same constructor, methods, state_dict layout (``g.`` / ``g._orig_mod.`` prefix) and error behaviour as
the reference's LightningModule; Lightning itself is optional (plain nn.Module shims otherwise).
The fast path for samplers is :meth:`denoise_positions`, which works on raw ``[N,3]`` tensors and a
cached :class:`jamun_b200.engine.Topology` -- no per-call graph cloning, no host synchronisation.
"""
from __future__ import annotations

import logging
import math
from typing import Callable, Dict, Optional, Tuple, Union

import numpy as np
import torch

from .. import _lib, engine, ops
from ..utils.unsqueeze_trailing import unsqueeze_trailing

try:  # pragma: no cover - Lightning is not installable in the build container
    import lightning.pytorch as pl

    _Base = pl.LightningModule
except Exception:  # noqa: BLE001
    pl = None

    class _Base(torch.nn.Module):
        """The members of LightningModule the path relies on (SURVEY 8b "Base-class shims")."""

        def __init__(self):
            super().__init__()
            self._logged: Dict[str, torch.Tensor] = {}
            self.hparams: Dict = {}

        @property
        def device(self) -> torch.device:
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, name, value, **kwargs):
            self._logged[name] = value

        def save_hyperparameters(self, *args, logger: bool = False, **kwargs):
            import inspect

            frame = inspect.currentframe().f_back
            init_args = {k: v for k, v in frame.f_locals.items() if k not in ("self", "__class__")}
            self.hparams = init_args


class _CompiledShim(torch.nn.Module):
    """Reproduces torch.compile's ``_orig_mod`` attribute so state_dict keys read ``g._orig_mod.*``
    (denoiser.py:37-42 of the reference); there is nothing to compile here."""

    def __init__(self, mod: torch.nn.Module):
        super().__init__()
        self._orig_mod = mod

    def forward(self, *args, **kwargs):
        return self._orig_mod(*args, **kwargs)


class SigmaContext:
    """Per-noise-level scalars, computed exactly as the reference does (fp32 tensor arithmetic)."""

    def __init__(self, sigma: float, average_squared_distance: float, max_radius: float, D: int = 3):
        s = torch.as_tensor(float(sigma), dtype=torch.float32)
        c_in, c_skip, c_out, c_noise = Denoiser.normalization_factors(s, average_squared_distance, D)
        r = torch.sqrt((max_radius ** 2) + 6 * (s ** 2)) / c_in
        self.sigma = float(s)
        self.sigma2 = float(s ** 2)
        self.c_in, self.c_skip, self.c_out, self.c_noise = float(c_in), float(c_skip), float(c_out), float(c_noise)
        self.r_cut = float(r)


class Denoiser(_Base):
    """The main denoiser model."""

    def __init__(
        self,
        arch: Callable[..., torch.nn.Module],
        optim: Callable[..., torch.optim.Optimizer],
        sigma_distribution: torch.distributions.Distribution,
        max_radius: float,
        average_squared_distance: float,
        add_fixed_noise: bool,
        add_fixed_ones: bool,
        align_noisy_input_during_training: bool,
        align_noisy_input_during_evaluation: bool,
        mean_center: bool,
        mirror_augmentation_rate: float,
        bond_loss_coefficient: float = 1.0,
        lr_scheduler_config: Optional[Dict] = None,
        use_torch_compile: bool = True,
        torch_compile_kwargs: Optional[Dict] = None,
    ):
        super().__init__()
        self.save_hyperparameters(logger=False)
        self.g = arch()
        if use_torch_compile:
            self.g = _CompiledShim(self.g)  # only selects the state-dict prefix
        py_logger = logging.getLogger("jamun")
        self.optim_factory = optim
        self.lr_scheduler_config = lr_scheduler_config
        self.sigma_distribution = sigma_distribution
        self.max_radius = max_radius
        self.add_fixed_noise = add_fixed_noise
        self.add_fixed_ones = add_fixed_ones
        if self.add_fixed_noise and self.add_fixed_ones:
            raise ValueError("Can't add fixed noise and fixed ones at the same time")
        self.average_squared_distance = average_squared_distance
        py_logger.info(f"Average squared distance = {self.average_squared_distance}")
        self.align_noisy_input_during_training = align_noisy_input_during_training
        self.align_noisy_input_during_evaluation = align_noisy_input_during_evaluation
        self.mean_center = mean_center
        self.mirror_augmentation_rate = mirror_augmentation_rate
        self.bond_loss_coefficient = bond_loss_coefficient
        self.max_num_neighbors: Optional[int] = 32  # torch_cluster default the reference inherits
        self._sigma_ctx: Dict[float, SigmaContext] = {}
        _lib.lib()  # fail loudly at construction if the CUDA library is absent

    # ------------------------------------------------------------------ helpers
    @property
    def arch_module(self) -> torch.nn.Module:
        return self.g._orig_mod if isinstance(self.g, _CompiledShim) else self.g

    def sigma_context(self, sigma) -> SigmaContext:
        key = float(torch.as_tensor(sigma, dtype=torch.float32))
        ctx = self._sigma_ctx.get(key)
        if ctx is None:
            if len(self._sigma_ctx) >= 16:  # training draws sigma from a continuous distribution: keep the cache bounded
                self._sigma_ctx.pop(next(iter(self._sigma_ctx)))
            ctx = self._sigma_ctx[key] = SigmaContext(key, self.average_squared_distance, self.max_radius)
        return ctx

    def topology_for(self, batch) -> engine.Topology:
        topo = batch["_topology"] if "_topology" in batch else None
        if topo is None or topo.device != batch.pos.device or topo.max_num_neighbors != self.max_num_neighbors:
            topo = engine.Topology(batch, batch.pos.device, self.max_num_neighbors)
            batch["_topology"] = topo
        return topo

    def denoise_positions(self, y: torch.Tensor, topo: engine.Topology, sigma, want_score: bool = True
                          ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """(xhat, score) for raw positions y [N,3] -- one fused pass, everything stays on the device."""
        ctx = self.sigma_context(sigma)
        plan = self.arch_module.plan(ctx.c_noise, y.device)
        y = y.contiguous()
        ybar, p = ops.center_scale(y, topo.chain_ptr, ctx.c_in, center=self.mean_center)
        topo.build_csr(ybar, ctx.r_cut)
        g = torch.empty_like(y)
        engine.e3conv_forward(plan, topo, p, ctx.r_cut, g)
        xhat = torch.empty_like(y)
        score = torch.empty_like(y) if want_score else None
        prm = _lib.WalkParams(c_in=ctx.c_in, c_skip=ctx.c_skip, c_out=ctx.c_out, sigma2=ctx.sigma2, delta=0.0, u=1.0,
                              a=0.0, z_sqrt_u=0.0, beta=1.0, clip=0.0, first=1, last=1, center=int(self.mean_center),
                              seed=0, step=0)
        v_dummy = torch.zeros_like(y)
        ops.walk_step(y, v_dummy, ybar, p, g, topo.chain_ptr, prm, None, xhat, score)
        if plan.gemm_kind == "f16" and not torch.cuda.is_current_stream_capturing() and topo.overflowed():
            engine.fall_back_to_tf32("a denoiser evaluation")
            return self.denoise_positions(y, topo, sigma, want_score)
        return xhat, score

    # ------------------------------------------------------------------ reference API
    def add_noise(self, x, sigma: Union[float, torch.Tensor]):
        if isinstance(sigma, torch.Tensor) and sigma.is_cuda:
            sigma = unsqueeze_trailing(sigma.to(x.pos), x.pos.ndim)
        else:  # host value: used as a Python scalar (a host-to-device copy of it would drain the stream)
            sigma = float(torch.as_tensor(sigma))
        y = x.clone("pos")
        if self.add_fixed_ones:
            noise = torch.ones_like(x.pos)
        elif self.add_fixed_noise:
            torch.manual_seed(0)
            num_batches = int(x.batch.max().item()) + 1
            if len(x.pos.shape) == 2:
                num_nodes_per_batch = x.pos.shape[0] // num_batches
                noise = torch.randn_like((x.pos[:num_nodes_per_batch])).repeat(num_batches, 1)
            if len(x.pos.shape) == 3:
                noise = torch.randn_like((x.pos[0])).repeat(num_batches, 1, 1)
        else:
            noise = torch.randn_like(x.pos)
        y.pos = x.pos + sigma * noise
        if torch.rand(()) < self.mirror_augmentation_rate:
            y.pos = -y.pos
        return y

    def score(self, y, sigma: Union[float, torch.Tensor]) -> torch.Tensor:
        """Compute the score function (x_hat(y) - y) / sigma^2 (y not mean-centred in the subtraction)."""
        topo = self.topology_for(y)
        _, score = self.denoise_positions(y.pos, topo, sigma, want_score=True)
        return score

    @classmethod
    def normalization_factors(cls, sigma, average_squared_distance: float, D: int = 3):
        """Normalization factors for the input and output."""
        sigma = torch.as_tensor(sigma)
        A = torch.as_tensor(average_squared_distance)
        B = torch.as_tensor(2 * D * sigma ** 2)
        c_in = 1.0 / torch.sqrt(A + B)
        c_skip = A / (A + B)
        c_out = torch.sqrt((A * B) / (A + B))
        c_noise = torch.log(sigma) / 4
        return c_in, c_skip, c_out, c_noise

    @classmethod
    def loss_weight(cls, sigma, average_squared_distance: float, D: int = 3):
        _, _, c_out, _ = cls.normalization_factors(sigma, average_squared_distance, D)
        return 1 / (c_out ** 2)

    def effective_radial_cutoff(self, sigma) -> torch.Tensor:
        sigma = torch.as_tensor(sigma)
        return torch.sqrt((self.max_radius ** 2) + 6 * (sigma ** 2))

    def add_edges(self, y, radial_cutoff: float, materialize: bool = False):
        """Radius graph + bonded edges as a receiver-sorted CSR held by the batch's Topology (K1).
        ``materialize=True`` additionally writes ``edge_index``/``bond_mask`` tensors (host sync)."""
        topo = self.topology_for(y)
        topo.build_csr(y.pos.contiguous(), float(radial_cutoff))
        # tag: the CSR held by `topo` was built from exactly this position tensor (E3Conv.forward checks it, so a clone whose
        # positions changed, or a topology rebuilt by another call, falls back to the explicit edge_index)
        y["_csr_ready"] = (topo.csr_generation, y.pos.data_ptr(), y.pos._version)
        if materialize:
            y.edge_index, y.bond_mask = topo.edge_index()
        return y

    def xhat_normalized(self, y, sigma: Union[float, torch.Tensor]):
        ctx = self.sigma_context(sigma)
        y = self.add_edges(y, ctx.r_cut)
        y_scaled = y.clone("pos")
        y_scaled.pos = y.pos * ctx.c_in
        y_scaled["_csr_ready"] = (y["_csr_ready"][0], "scaled")  # same graph, positions scaled by c_in (denoiser.py:192)
        xhat = y.clone("pos")
        c_noise = torch.tensor([ctx.c_noise], dtype=torch.float32)
        g_pred = self.g(y_scaled, c_noise, ctx.r_cut)
        xhat.pos = ctx.c_skip * y.pos + ctx.c_out * g_pred.pos
        return xhat

    def xhat(self, y, sigma: Union[float, torch.Tensor]):
        """Compute the denoised prediction (graph in, graph out)."""
        topo = self.topology_for(y)
        xh, _ = self.denoise_positions(y.pos, topo, sigma, want_score=False)
        out = y.clone("pos")
        out.pos = xh
        return out

    def noise_and_denoise(self, x, sigma, align_noisy_input: bool):
        from ..utils import align_A_to_B_batched, mean_center

        with torch.no_grad():
            if self.mean_center:
                x = mean_center(x)
            y = self.add_noise(x, sigma)  # `sigma` stays where it is (host in training_step: no read-back, no copy)
            if self.mean_center:
                y = mean_center(y)
            if align_noisy_input:
                y = align_A_to_B_batched(y, x)
        xhat = self.xhat_with_grad(y, sigma) if (torch.is_grad_enabled() and self.training) else self.xhat(y, sigma)
        return xhat, y

    def xhat_with_grad(self, y, sigma: Union[float, torch.Tensor]):
        """xhat with an autograd graph to the parameters: forward and backward on this library's kernels through the
        torch.library operators of jamun_b200/autograd_ops.py (composition in jamun_b200/training.py)."""
        from .. import training

        topo = self.topology_for(y)
        out = y.clone("pos")
        out.pos = training.xhat_positions(self, y.pos, topo, sigma)
        return out

    def compute_loss(self, x, xhat, sigma):
        """Per-graph loss (differentiable w.r.t. xhat.pos; denoiser.py:251-287) -- one kernel forward, one backward."""
        from .. import autograd_ops  # noqa: F401
        from ..utils import mean_center

        if self.mean_center:
            x = mean_center(x)
        D = xhat.pos.shape[-1]
        topo = self.topology_for(x)
        lw = x.loss_weight.to(xhat.pos.device, torch.float32).contiguous() if "loss_weight" in x else None
        scale = float(self.loss_weight(torch.as_tensor(float(sigma)), self.average_squared_distance, D))
        loss, raw_g, rmsd_g = torch.ops.jamun_b200.coordinate_loss(xhat.pos, x.pos.contiguous(), topo.chain_of, topo.chain_ptr, lw,
                                                                   scale, float(torch.as_tensor(sigma)))
        return loss, {"coordinate_loss": loss, "raw_coordinate_loss": raw_g, "scaled_rmsd": rmsd_g}

    def noise_and_compute_loss(self, x, sigma, align_noisy_input: bool):
        xhat, _ = self.noise_and_denoise(x, sigma, align_noisy_input=align_noisy_input)
        return self.compute_loss(x, xhat, sigma)

    def _to_device_scalar(self, t: torch.Tensor) -> torch.Tensor:
        """`t.to(self.device)` for a 0-dim host tensor as a fill kernel: a pageable host-to-device copy would wait for the stream."""
        if t.is_cuda or t.ndim != 0 or torch.device(self.device).type != "cuda":
            return t.to(self.device)
        return torch.full((), float(t), dtype=t.dtype, device=self.device)

    def training_step(self, batch, batch_idx: int):
        """denoiser.py:299-319.  Forward and backward run on this library's kernels (jamun_b200/training.py)."""
        sigma_host = self.sigma_distribution.sample()  # the distributions live on the host: the noise level is known there ...
        sigma = self._to_device_scalar(sigma_host)
        # ... so the per-sigma scalars (c_in, c_skip, c_out, cut-off, loss weight) are taken from the host copy -- reading them
        # back from `sigma` would drain the stream at the start of every step and stop the host from running ahead
        loss, aux = self.noise_and_compute_loss(batch, sigma if sigma_host.is_cuda else sigma_host,
                                                align_noisy_input=self.align_noisy_input_during_training)
        aux["loss"] = loss
        for key in aux:
            aux[key] = aux[key].mean()
            self.log(f"train/{key}", aux[key], prog_bar=False, batch_size=batch.num_graphs, sync_dist=False)
        return {"sigma": sigma, **aux}

    def validation_step(self, batch, batch_idx: int):
        sigma_host = self.sigma_distribution.sample()
        sigma = self._to_device_scalar(sigma_host)
        loss, aux = self.noise_and_compute_loss(batch, sigma if sigma_host.is_cuda else sigma_host,
                                                align_noisy_input=self.align_noisy_input_during_training)
        aux["loss"] = loss
        for key in aux:
            aux[key] = aux[key].mean()
            self.log(f"val/{key}", aux[key], prog_bar=(key == "scaled_rmsd"), batch_size=batch.num_graphs, sync_dist=True)
        return {"sigma": sigma, **aux}

    def configure_optimizers(self):
        optimizer = self.optim_factory(params=self.parameters())
        out = {"optimizer": optimizer}
        if self.lr_scheduler_config:
            scheduler = self.lr_scheduler_config.pop("scheduler")
            out["lr_scheduler"] = {"scheduler": scheduler(optimizer), **self.lr_scheduler_config}
        return out

    # ------------------------------------------------------------------ checkpoints
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts ``g.`` and ``g._orig_mod.`` prefixes and ignores e3nn's constant buffers."""
        own = set(self.state_dict().keys())
        want_shim = isinstance(self.g, _CompiledShim)
        fixed = {}
        for k, v in state_dict.items():
            k2 = k
            if k.startswith("g._orig_mod.") and not want_shim:
                k2 = "g." + k[len("g._orig_mod."):]
            elif k.startswith("g.") and not k.startswith("g._orig_mod.") and want_shim:
                k2 = "g._orig_mod." + k[2:]
            if k2 not in own:
                # e3nn's constant buffers: empty weight/bias of tensor products with external weights,
                # output masks, Wigner-3j tables of the compiled tensor products
                if v.numel() == 0 or k2.endswith("output_mask") or "_w3j" in k2 or "_compiled_main" in k2 \
                        or k2.endswith(".sh._lmax"):
                    continue
                if ".tp." in k2:  # anything else under a tensor product would be a learnable tensor we do not model
                    import warnings

                    warnings.warn(f"load_state_dict: ignoring unknown tensor-product entry {k2!r} with {v.numel()} elements "
                                  f"(only e3nn's constant buffers are expected there)", stacklevel=2)
                    continue
            fixed[k2] = v
        return super().load_state_dict(fixed, strict=strict, **kw)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path: str, map_location=None, **overrides):
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        hparams = dict(ckpt.get("hyper_parameters", {}))
        hparams.update(overrides)
        model = cls(**hparams)
        model.load_state_dict(ckpt["state_dict"], strict=True)
        return model
