"""Synthetic peptide-shaped inputs for benchmarks and parity tests.

Per chain of n atoms: coordinates (nm) from a branched random walk -- atom k is placed 0.15 nm
from a uniformly chosen one of the previous <=3 atoms, rejected if closer than 0.11 nm to any
earlier atom -- then mean-centred; rng = numpy.random.default_rng(1234 + chain_id).  Topology
indices are drawn with the marginals of real heavy-atom peptides; bonded edges are the walk's
parent->child pairs, one direction only (as /root/reference/src/jamun/data/_mdtraj.py:73).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

# heavy-atom counts of the 20 residues (backbone 4 + side chain), SURVEY 8(d)
_RES_HEAVY = [4, 5, 6, 6, 7, 7, 7, 8, 8, 8, 8, 8, 9, 9, 9, 10, 11, 11, 12, 14]


def peptide_sizes(num_chains: int, residues_per_chain: int, seed: int = 0) -> List[int]:
    """Heavy-atom counts of uncapped peptides: sum of residue sizes + 1 (OXT)."""
    rng = np.random.default_rng(seed)
    picks = rng.integers(0, len(_RES_HEAVY), size=(num_chains, residues_per_chain))
    return [int(sum(_RES_HEAVY[i] for i in row) + 1) for row in picks]


def make_chain(n: int, chain_id: int, n_res: int = 2) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(1234 + chain_id)
    pos = np.zeros((n, 3), dtype=np.float64)
    parent = np.zeros(n, dtype=np.int64)
    k = 1
    while k < n:
        p = int(rng.integers(max(0, k - 3), k))
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        cand = pos[p] + 0.15 * d
        if np.min(np.linalg.norm(pos[:k] - cand, axis=1)) < 0.11:
            continue
        pos[k] = cand
        parent[k] = p
        k += 1
    pos -= pos.mean(axis=0, keepdims=True)
    atom_type = rng.choice([0, 1, 2, 4], size=n, p=[0.62, 0.19, 0.17, 0.02])
    atom_code = rng.integers(0, 7, size=n)
    res_code = rng.integers(0, 20, size=n)
    res_seq = (np.arange(n) * n_res) // n
    bonds = np.stack([parent[1:], np.arange(1, n)]) if n > 1 else np.zeros((2, 0), dtype=np.int64)
    return dict(pos=pos.astype(np.float32), atom_type_index=atom_type.astype(np.int32),
                atom_code_index=atom_code.astype(np.int32), residue_code_index=res_code.astype(np.int32),
                residue_sequence_index=res_seq.astype(np.int32), edge_index=bonds.astype(np.int64))


def make_tensors(sizes: Sequence[int], n_res: int = 2, first_chain_id: int = 0) -> Dict[str, torch.Tensor]:
    """Concatenate chains into batch-level tensors (the members of a PyG Batch the path reads)."""
    chains = [make_chain(n, first_chain_id + c, n_res) for c, n in enumerate(sizes)]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    out = {
        "pos": torch.from_numpy(np.concatenate([c["pos"] for c in chains])),
        "batch": torch.from_numpy(np.repeat(np.arange(len(sizes)), sizes)).long(),
        "edge_index": torch.from_numpy(np.concatenate([c["edge_index"] + offs[i] for i, c in enumerate(chains)], axis=1)).long(),
        "ptr": torch.from_numpy(offs).long(),
        "num_graphs": len(sizes),
    }
    for key in ("atom_type_index", "atom_code_index", "residue_code_index", "residue_sequence_index"):
        out[key] = torch.from_numpy(np.concatenate([c[key] for c in chains]))
    out["loss_weight"] = torch.ones(len(sizes))
    return out


def workload_sizes(name: str, num_chains: int) -> List[int]:
    """Named BASELINE configs -> per-chain atom counts."""
    if name == "ala2_capped":  # C1: 22 atoms
        return [22] * num_chains
    if name == "2AA":  # C2: uncapped 2AA, 9..29 heavy atoms
        return peptide_sizes(num_chains, 2, seed=2)
    if name == "4AA":  # C3: uncapped 4AA, 17..57 heavy atoms
        return peptide_sizes(num_chains, 4, seed=4)
    if name == "protein1000":  # C4
        return [1000] * num_chains
    raise ValueError(name)
