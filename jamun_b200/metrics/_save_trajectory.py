"""SaveTrajectory (mirror of /root/reference/src/jamun/metrics/_save_trajectory.py:12-97 and metrics/_utils.py:31-110): accumulates
the per-chain sample trajectories of one dataset and writes them in the reference's on-disk layout

    sampler/<label>/topology.pdb
    sampler/<label>/predicted_samples/{npy,pdb,dcd}/{<trajectory index>,joined}.<ext>

npy arrays have shape [atoms, frames, 3] in nm (``joined``: all chains' frames concatenated along the frame axis).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Union

import numpy as np
import torch

from .. import pdb as _pdb


class SaveTrajectory:
    def __init__(self, label: str, topology: "_pdb.Topology", sample_key: str = "xhat_traj", output_root: str = ".",
                 init_positions_nm: Optional[np.ndarray] = None):
        self.label, self.topology, self.sample_key = label, topology, sample_key
        self.output_dir = os.path.join(output_root, "sampler", label)
        self.pred_samples_dir = os.path.join(self.output_dir, "predicted_samples")
        self.pred_samples_extensions = ["npy", "pdb", "dcd"]
        for ext in self.pred_samples_extensions:
            os.makedirs(os.path.join(self.pred_samples_dir, ext), exist_ok=True)
        self.samples: List[torch.Tensor] = []  # each [atoms, frames, 3]
        self.num_chains_seen = 0
        self.init_positions_nm = init_positions_nm

    def filename_pred(self, trajectory_index: Union[int, str], extension: str) -> str:
        if extension not in self.pred_samples_extensions:
            raise ValueError(f"Invalid extension: {extension}")
        return os.path.join(self.pred_samples_dir, extension, f"{trajectory_index}.{extension}")

    # ---- metric protocol
    def on_sample_start(self) -> None:
        if self.init_positions_nm is not None:
            _pdb.write_pdb(os.path.join(self.output_dir, "topology.pdb"), self.topology, self.init_positions_nm[None])

    def update(self, sample) -> None:
        if "dataset_label" in sample and sample["dataset_label"] not in (None, self.label):
            raise ValueError(f"Sample dataset label {sample['dataset_label']} does not match expected label {self.label}.")
        s = sample[self.sample_key]
        if s.ndim != 3:
            raise ValueError(f"Invalid sample shape: {s.shape}, expected (num_atoms, num_frames, 3).")
        if s.shape[0] != self.topology.n_atoms:
            raise ValueError(f"{s.shape[0]} atoms in the sample, {self.topology.n_atoms} in the topology")
        self.samples.append(s.detach().cpu())

    def on_after_sample_batch(self) -> None:
        self.compute()
        self.num_chains_seen = len(self.samples)

    def on_sample_end(self) -> None:
        pass

    def sample_tensors(self, *, new: bool) -> List[torch.Tensor]:
        return self.samples[self.num_chains_seen:] if new else self.samples

    def joined_sample_tensor(self) -> torch.Tensor:
        return torch.cat(self.samples, dim=1)  # atoms, (chains frames), 3

    def compute(self) -> Dict[str, float]:
        for k, s in enumerate(self.sample_tensors(new=True), start=self.num_chains_seen):
            arr = s.numpy()
            np.save(self.filename_pred(k, "npy"), arr)
            frames = np.transpose(arr, (1, 0, 2))
            _pdb.write_pdb(self.filename_pred(k, "pdb"), self.topology, frames)
            _pdb.write_dcd(self.filename_pred(k, "dcd"), frames)
        if self.samples:
            joined = self.joined_sample_tensor().numpy()
            np.save(self.filename_pred("joined", "npy"), joined)
            frames = np.transpose(joined, (1, 0, 2))
            _pdb.write_pdb(self.filename_pred("joined", "pdb"), self.topology, frames)
            _pdb.write_dcd(self.filename_pred("joined", "dcd"), frames)
        return {}
