"""E3Conv (mirror of /root/reference/src/jamun/model/arch/e3conv.py:12-138): same constructor, same
state_dict layout, forward executed by the sm_100a kernel sequence in jamun_b200.engine."""
from __future__ import annotations

from typing import Callable, Union

import torch

from ... import engine
from ...irreps import Irreps
from ..atom_embedding import AtomEmbeddingWithResidueInformation, SimpleAtomEmbedding
from ..noise_conditioning import NoiseConditionalScaling, NoiseConditionalSkipConnection


class E3Conv(torch.nn.Module):
    def __init__(self, irreps_out: Union[str, Irreps], irreps_hidden: Union[str, Irreps], irreps_sh: Union[str, Irreps],
                 hidden_layer_factory: Callable[..., torch.nn.Module], output_head_factory: Callable[..., torch.nn.Module],
                 use_residue_information: bool, n_layers: int, edge_attr_dim: int, atom_type_embedding_dim: int,
                 atom_code_embedding_dim: int, residue_code_embedding_dim: int, residue_index_embedding_dim: int,
                 use_residue_sequence_index: bool, test_equivariance: bool = False):
        super().__init__()
        self.test_equivariance = test_equivariance
        self.irreps_out, self.irreps_hidden, self.irreps_sh = Irreps(irreps_out), Irreps(irreps_hidden), Irreps(irreps_sh)
        self.n_layers, self.edge_attr_dim = n_layers, edge_attr_dim
        self.bonded_edge_attr_dim, self.radial_edge_attr_dim = edge_attr_dim // 2, (edge_attr_dim + 1) // 2
        self.embed_bondedness = torch.nn.Embedding(2, self.bonded_edge_attr_dim)
        if use_residue_information:
            self.atom_embedder = AtomEmbeddingWithResidueInformation(
                atom_type_embedding_dim, atom_code_embedding_dim, residue_code_embedding_dim,
                residue_index_embedding_dim, use_residue_sequence_index)
        else:
            self.atom_embedder = SimpleAtomEmbedding(atom_type_embedding_dim + atom_code_embedding_dim
                                                     + residue_code_embedding_dim + residue_index_embedding_dim)
        self.initial_noise_scaling = NoiseConditionalScaling(self.atom_embedder.irreps_out)
        self.initial_projector = hidden_layer_factory(irreps_in=self.initial_noise_scaling.irreps_out,
                                                      irreps_out=self.irreps_hidden, irreps_sh=self.irreps_sh,
                                                      edge_attr_dim=edge_attr_dim)
        self.layers = torch.nn.ModuleList()
        self.noise_scalings = torch.nn.ModuleList()
        self.skip_connections = torch.nn.ModuleList()
        for _ in range(n_layers):
            self.layers.append(hidden_layer_factory(irreps_in=self.irreps_hidden, irreps_out=self.irreps_hidden,
                                                    irreps_sh=self.irreps_sh, edge_attr_dim=self.edge_attr_dim))
            self.noise_scalings.append(NoiseConditionalScaling(self.irreps_hidden))
            self.skip_connections.append(NoiseConditionalSkipConnection(self.irreps_hidden))
        self.output_head = output_head_factory(irreps_in=self.irreps_hidden, irreps_out=self.irreps_out)
        self.output_gain = torch.nn.Parameter(torch.tensor(0.0))
        self._plan = None
        self._plan_key = None

    # ---- plan cache: re-pack when a parameter changed in place or the noise level changed
    def plan(self, c_noise: float, device) -> "engine.E3ConvPlan":
        import os

        # everything but the noise level: a change here rebuilds the packed weight images; a new noise level alone only
        # re-evaluates the noise-conditioning MLPs (validation / training draw sigma from a continuous distribution)
        key = (str(device), tuple(p._version for p in self.parameters()),
               tuple(p.data_ptr() for p in self.parameters()), os.environ.get("JAMUN_B200_GEMM", engine.GEMM_KIND))
        if self._plan is None or self._plan_key != key:
            self._plan = engine.E3ConvPlan(self, float(c_noise), device)
            self._plan_key = key
        elif self._plan.c_noise != float(c_noise):
            self._plan.set_noise(self, float(c_noise))
        return self._plan

    def forward(self, data, c_noise: torch.Tensor, effective_radial_cutoff: float):
        pos = data["pos"]
        if not pos.is_cuda:
            raise RuntimeError("jamun_b200.E3Conv runs on CUDA only (no CPU fallback)")
        topo = data["_topology"] if "_topology" in data else None
        if topo is None:
            topo = engine.Topology(data, pos.device, max_num_neighbors=None)
            data["_topology"] = topo
        tag = data["_csr_ready"] if "_csr_ready" in data else None
        if tag != (topo.csr_generation, pos.data_ptr(), pos._version) and tag != (topo.csr_generation, "scaled"):
            # no CSR, or one built for other positions / overwritten since: use the edge list the caller supplied
            if "bond_mask" not in data:
                raise RuntimeError("E3Conv.forward: the graph carries neither a current CSR (Denoiser.add_edges on these "
                                   "positions) nor an explicit edge_index + bond_mask")
            topo.set_csr_from_edge_index(data["edge_index"], data["bond_mask"])
        plan = self.plan(float(torch.as_tensor(c_noise).reshape(-1)[0]), pos.device)
        g = torch.empty_like(pos)
        engine.e3conv_forward(plan, topo, pos.contiguous(), float(effective_radial_cutoff), g)
        if plan.gemm_kind == "f16" and not torch.cuda.is_current_stream_capturing() and topo.overflowed():
            engine.fall_back_to_tf32("E3Conv.forward")
            plan = self.plan(float(torch.as_tensor(c_noise).reshape(-1)[0]), pos.device)
            engine.e3conv_forward(plan, topo, pos.contiguous(), float(effective_radial_cutoff), g)
        data["pos"] = g
        return data
