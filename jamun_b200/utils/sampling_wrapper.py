"""ModelSamplingWrapper (mirror of /root/reference/src/jamun/utils/sampling_wrapper.py:9-83)."""
from __future__ import annotations

from typing import Dict, List

import torch

from .. import ops


class ModelSamplingWrapper:
    """Wrapper to sample positions from a model."""

    def __init__(self, model, init_graphs, sigma: float):
        self._model = model
        self.init_graphs = init_graphs
        self.sigma = sigma
        self._topology = None

    @property
    def device(self) -> torch.device:
        return self._model.device

    @property
    def topology(self):
        if self._topology is None:
            self._topology = self._model.topology_for(self.init_graphs)
        return self._topology

    def sample_initial_noisy_positions(self, noise: torch.Tensor = None) -> torch.Tensor:
        from ..sampling.mcmc.functional._splitting import _call_counter, _philox_seed

        pos = self.init_graphs.pos.contiguous()
        out = torch.empty_like(pos)
        return ops.gaussian_axpy(pos, 1.0, float(self.sigma), noise, _philox_seed(), next(_call_counter) << 32, out)

    def __getattr__(self, name):
        return getattr(self._model, name)

    def score(self, y, sigma, *args, **kwargs):
        self._check(y)
        return self._model.denoise_positions(y, self.topology, sigma, want_score=True)[1]

    def xhat(self, y, sigma, *args, **kwargs):
        self._check(y)
        return self._model.denoise_positions(y, self.topology, sigma, want_score=False)[0]

    def fused_walk(self, mcmc, y_init, sigma, **kwargs):
        from ..sampling.mcmc.functional import fused_baoab

        self._check(y_init)
        return fused_baoab(self._model, self.topology, y_init, sigma, **mcmc.params(**kwargs))

    def _check(self, positions: torch.Tensor):
        assert len(positions) == self.init_graphs.num_nodes, "The number of positions and nodes should be the same"
        assert positions.shape[1] == 3, "Positions tensor should have a shape of (n, 3)"

    def positions_to_graph(self, positions: torch.Tensor):
        """Wraps a tensor of positions to a graph with these positions as an attribute."""
        self._check(positions)
        input_graphs = self.init_graphs.clone()
        input_graphs.pos = positions
        self.input_graphs = input_graphs
        return input_graphs.to(positions.device)

    def unbatch_samples(self, samples: Dict[str, torch.Tensor]) -> List:
        """Unbatch samples: per graph, tensors of shape [atoms, (frames,) 3]."""
        if "batch" not in self.init_graphs:
            raise ValueError("The initial graph does not have a batch attribute.")
        output_graphs = self.init_graphs.clone().to_data_list()
        ptr = self.init_graphs.ptr.tolist()
        for key, value in samples.items():
            if value is None or not isinstance(value, torch.Tensor) or value.ndim not in [2, 3]:
                continue
            if value.ndim == 3:
                value = value.permute(1, 0, 2)  # num_frames atoms coords -> atoms num_frames coords
            for g, output_graph in enumerate(output_graphs):
                if key in output_graph:
                    raise ValueError(f"Key {key} already exists in the output graph.")
                chunk = value[ptr[g]:ptr[g + 1]]
                if chunk.shape[0] != output_graph.num_nodes:
                    raise ValueError(f"Number of nodes in unbatched value ({chunk.shape[0]}) for key {key} does not match "
                                     f"number of nodes in output graph ({output_graph.num_nodes}).")
                output_graph[key] = chunk
        return output_graphs
