"""Hydra-free construction of the default model (what configs/experiment/*.yaml + hydra_config/model/arch/e3conv.yaml
compose in the reference; same values as /root/reference/profiling/standalone_training.py:48-101)."""
from __future__ import annotations

import functools

import torch


def default_arch(n_layers: int = 5):
    from . import e3tools, model

    return functools.partial(
        model.arch.E3Conv,
        irreps_out="1x1e", irreps_hidden="120x0e + 32x1e", irreps_sh="1x0e + 1x1e", n_layers=n_layers, edge_attr_dim=64,
        atom_type_embedding_dim=8, atom_code_embedding_dim=8, residue_code_embedding_dim=32,
        residue_index_embedding_dim=8, use_residue_information=True, use_residue_sequence_index=False,
        hidden_layer_factory=functools.partial(e3tools.nn.ConvBlock, conv=e3tools.nn.Conv),
        output_head_factory=functools.partial(e3tools.nn.EquivariantMLP, irreps_hidden_list=["120x0e + 32x1e"]),
    )


def default_denoiser(sigma: float = 0.04, max_radius: float = 1.0, average_squared_distance: float = 0.332,
                     use_torch_compile: bool = False, n_layers: int = 5, **kw):
    from . import distributions, model

    args = dict(arch=default_arch(n_layers), optim=functools.partial(torch.optim.Adam, lr=2e-3),
                sigma_distribution=distributions.ConstantSigma(sigma), max_radius=max_radius,
                average_squared_distance=average_squared_distance, add_fixed_noise=False, add_fixed_ones=False,
                align_noisy_input_during_training=True, align_noisy_input_during_evaluation=True, mean_center=True,
                mirror_augmentation_rate=0.0, use_torch_compile=use_torch_compile)
    args.update(kw)
    return model.Denoiser(**args)
