"""Duck-typed stand-ins for torch_geometric.data.{Data,Batch} / jamun.utils.DataWithResidueInformation
with exactly the members the walk-jump path touches (SURVEY 8b "Data container";
/root/reference/src/jamun/utils/data_with_residue_info.py:5-33)."""
from __future__ import annotations

import copy
from typing import Any, Dict, Iterable, List, Optional

import torch

_NODE_KEYS = ("pos", "atom_type_index", "atom_code_index", "residue_code_index", "residue_sequence_index")


class Data:
    def __init__(self, **kwargs: Any):
        self.__dict__["_store"] = dict(kwargs)

    # attribute and item access by name
    def __getattr__(self, key):
        store = self.__dict__.get("_store", {})
        if key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        self._store[key] = value

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store and self._store[key] is not None

    def keys(self):
        return list(self._store.keys())

    @property
    def num_nodes(self) -> int:
        for k in _NODE_KEYS:
            v = self._store.get(k)
            if isinstance(v, torch.Tensor):
                return v.shape[0]
        raise ValueError("cannot infer num_nodes")

    def clone(self, *keys: str):
        """PyG semantics: clone() deep-copies tensors; clone('pos') shallow-copies and clones only those keys."""
        out = self.__class__.__new__(self.__class__)
        out.__dict__["_store"] = dict(self._store)
        for k, v in self._store.items():
            if isinstance(v, torch.Tensor) and (not keys or k in keys):
                out._store[k] = v.clone()
        return out

    def to(self, device):
        out = self.__class__.__new__(self.__class__)
        out.__dict__["_store"] = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self._store.items()}
        return out


class DataWithResidueInformation(Data):
    pass


class Batch(DataWithResidueInformation):
    @property
    def num_graphs(self) -> int:
        return int(self._store["ptr"].numel() - 1)

    @classmethod
    def from_data_list(cls, data_list: Iterable[Data]) -> "Batch":
        data_list = list(data_list)
        sizes = [d.num_nodes for d in data_list]
        ptr = torch.zeros(len(sizes) + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(torch.tensor(sizes), 0)
        store: Dict[str, Any] = {}
        for k in _NODE_KEYS:
            if all(k in d for d in data_list):
                store[k] = torch.cat([d[k] for d in data_list], dim=0)
        store["edge_index"] = torch.cat([d["edge_index"] + int(ptr[i]) for i, d in enumerate(data_list)], dim=1)
        store["batch"] = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
        store["ptr"] = ptr
        lw = [d["loss_weight"] if "loss_weight" in d else 1.0 for d in data_list]
        store["loss_weight"] = torch.as_tensor(lw, dtype=torch.float32)
        store["dataset_label"] = [d["dataset_label"] if "dataset_label" in d else None for d in data_list]
        out = cls.__new__(cls)
        out.__dict__["_store"] = store
        return out

    @classmethod
    def from_tensors(cls, tensors: Dict[str, Any]) -> "Batch":
        """From jamun_b200.synthetic.make_tensors output."""
        out = cls.__new__(cls)
        store = {k: v for k, v in tensors.items() if k != "num_graphs"}
        out.__dict__["_store"] = store
        return out

    def to_data_list(self) -> List[Data]:
        ptr = self._store["ptr"].tolist()
        out = []
        ei = self._store["edge_index"]
        for g in range(len(ptr) - 1):
            lo, hi = ptr[g], ptr[g + 1]
            d = DataWithResidueInformation()
            for k in _NODE_KEYS:
                if k in self:
                    d[k] = self._store[k][lo:hi]
            m = (ei[1] >= lo) & (ei[1] < hi)
            d["edge_index"] = ei[:, m] - lo
            if "loss_weight" in self:
                d["loss_weight"] = self._store["loss_weight"][g]
            if "dataset_label" in self and self._store["dataset_label"] is not None:
                d["dataset_label"] = self._store["dataset_label"][g]
            out.append(d)
        return out
