"""Atom embeddings (mirror of /root/reference/src/jamun/model/atom_embedding.py:22-76); evaluated by jamun_atom_embed."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..irreps import Irreps


class SimpleAtomEmbedding(nn.Module):
    def __init__(self, embedding_dim: int, max_value: int = 20):
        super().__init__()
        self.embedding = nn.Embedding(max_value, embedding_dim)
        self.irreps_out = Irreps(f"{embedding_dim}x0e")


class AtomEmbeddingWithResidueInformation(nn.Module):
    def __init__(self, atom_type_embedding_dim: int, atom_code_embedding_dim: int, residue_code_embedding_dim: int,
                 residue_index_embedding_dim: int, use_residue_sequence_index: bool, num_atom_types: int = 20,
                 max_sequence_length: int = 10, num_atom_codes: int = 10, num_residue_types: int = 25):
        super().__init__()
        self.atom_type_embedding = nn.Embedding(num_atom_types, atom_type_embedding_dim)
        self.atom_code_embedding = nn.Embedding(num_atom_codes, atom_code_embedding_dim)
        self.residue_code_embedding = nn.Embedding(num_residue_types, residue_code_embedding_dim)
        self.residue_index_embedding = nn.Embedding(max_sequence_length, residue_index_embedding_dim)
        self.use_residue_sequence_index = use_residue_sequence_index
        # the reference repeats atom_type_embedding_dim for the atom-code block (atom_embedding.py:54-56)
        self.irreps_out = Irreps(f"{atom_type_embedding_dim}x0e + {atom_type_embedding_dim}x0e + "
                                 f"{residue_code_embedding_dim}x0e + {residue_index_embedding_dim}x0e")

    def forward(self, data):
        """Module-level compatibility forward on jamun_atom_embed."""
        from .. import ops

        dev = self.atom_type_embedding.weight.device
        idx = [data[k].to(dev, torch.int32).contiguous() for k in ("atom_type_index", "atom_code_index", "residue_code_index")]
        idx.append(data["residue_sequence_index"].to(dev, torch.int32).contiguous() if self.use_residue_sequence_index else None)
        return ops.atom_embed(idx, [t.detach().contiguous() for t in self.tables()], None)

    def tables(self):
        return [self.atom_type_embedding.weight, self.atom_code_embedding.weight, self.residue_code_embedding.weight,
                self.residue_index_embedding.weight]
