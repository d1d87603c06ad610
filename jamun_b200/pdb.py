"""Topologies in, trajectories out without mdtraj: a heavy-atom PDB reader producing the graph tensors of
``preprocess_topology`` (/root/reference/src/jamun/data/_mdtraj.py:56-89), and PDB / DCD / npy writers for the on-disk layout
of ``SaveTrajectory`` (/root/reference/src/jamun/metrics/_save_trajectory.py:17-41,78-97).

Bonds come from per-residue heavy-atom templates plus the peptide bond C(i)-N(i+1), as mdtraj's ``create_standard_bonds``
derives them for standard residues; CONECT records are honoured for anything else.  Bonded edges are single-direction
(atom1 -> atom2 in file order), as the reference's ``edge_index`` (``_mdtraj.py:73``).  Coordinates are nm in memory
(mdtraj's unit) and Angstrom on disk.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .data import DataWithResidueInformation
from .utils.residue_metadata import ResidueMetadata, encode_atom_code, encode_atom_type, encode_residue

_BB = [("N", "CA"), ("CA", "C"), ("C", "O"), ("C", "OXT")]
_SIDE: Dict[str, List[Tuple[str, str]]] = {
    "GLY": [], "ALA": [("CA", "CB")], "SER": [("CA", "CB"), ("CB", "OG")], "CYS": [("CA", "CB"), ("CB", "SG")],
    "VAL": [("CA", "CB"), ("CB", "CG1"), ("CB", "CG2")], "THR": [("CA", "CB"), ("CB", "OG1"), ("CB", "CG2")],
    "PRO": [("CA", "CB"), ("CB", "CG"), ("CG", "CD"), ("CD", "N")],
    "ILE": [("CA", "CB"), ("CB", "CG1"), ("CB", "CG2"), ("CG1", "CD1")],
    "LEU": [("CA", "CB"), ("CB", "CG"), ("CG", "CD1"), ("CG", "CD2")],
    "ASP": [("CA", "CB"), ("CB", "CG"), ("CG", "OD1"), ("CG", "OD2")], "ASN": [("CA", "CB"), ("CB", "CG"), ("CG", "OD1"), ("CG", "ND2")],
    "MET": [("CA", "CB"), ("CB", "CG"), ("CG", "SD"), ("SD", "CE")],
    "GLU": [("CA", "CB"), ("CB", "CG"), ("CG", "CD"), ("CD", "OE1"), ("CD", "OE2")],
    "GLN": [("CA", "CB"), ("CB", "CG"), ("CG", "CD"), ("CD", "OE1"), ("CD", "NE2")],
    "LYS": [("CA", "CB"), ("CB", "CG"), ("CG", "CD"), ("CD", "CE"), ("CE", "NZ")],
    "HIS": [("CA", "CB"), ("CB", "CG"), ("CG", "ND1"), ("CG", "CD2"), ("ND1", "CE1"), ("CD2", "NE2"), ("CE1", "NE2")],
    "PHE": [("CA", "CB"), ("CB", "CG"), ("CG", "CD1"), ("CG", "CD2"), ("CD1", "CE1"), ("CD2", "CE2"), ("CE1", "CZ"), ("CE2", "CZ")],
    "ARG": [("CA", "CB"), ("CB", "CG"), ("CG", "CD"), ("CD", "NE"), ("NE", "CZ"), ("CZ", "NH1"), ("CZ", "NH2")],
    "TYR": [("CA", "CB"), ("CB", "CG"), ("CG", "CD1"), ("CG", "CD2"), ("CD1", "CE1"), ("CD2", "CE2"), ("CE1", "CZ"), ("CE2", "CZ"),
            ("CZ", "OH")],
    "TRP": [("CA", "CB"), ("CB", "CG"), ("CG", "CD1"), ("CG", "CD2"), ("CD1", "NE1"), ("NE1", "CE2"), ("CD2", "CE2"), ("CD2", "CE3"),
            ("CE2", "CZ2"), ("CE3", "CZ3"), ("CZ2", "CH2"), ("CZ3", "CH2")],
    "ACE": [("CH3", "C"), ("C", "O")], "NME": [("N", "C"), ("N", "CH3")],
}
_PROTEIN = set(_SIDE)


@dataclass
class Atom:
    serial: int
    name: str
    element: str
    res_name: str
    res_seq: int
    chain: str
    residue_index: int = 0


@dataclass
class Topology:
    """The slice of an mdtraj.Topology this path needs: heavy protein atoms, residues, bonds."""

    atoms: List[Atom] = field(default_factory=list)
    bonds: List[Tuple[int, int]] = field(default_factory=list)

    @property
    def n_atoms(self) -> int:
        return len(self.atoms)


def _element(name: str, column: str) -> str:
    e = column.strip().capitalize()
    if e:
        return e
    letters = "".join(ch for ch in name if ch.isalpha())
    return letters[:1].upper()


def read_pdb(path_or_text: str) -> Tuple[Topology, np.ndarray]:
    """Heavy protein atoms of the first MODEL -> (Topology, positions [n, 3] in nm)."""
    text = path_or_text if "\n" in path_or_text else open(path_or_text).read()
    atoms, xyz, serial_to_idx, conect = [], [], {}, []
    res_keys: Dict[Tuple[str, int, str], int] = {}
    for line in text.splitlines():
        rec = line[:6].strip()
        if rec == "ENDMDL":
            break
        if rec in ("ATOM", "HETATM"):
            name, res_name = line[12:16].strip(), line[17:20].strip()
            element = _element(name, line[76:78] if len(line) >= 78 else "")
            if res_name not in _PROTEIN or element == "H" or (not line[76:78].strip() and name[:1] == "H"):
                continue  # mdtraj selection "protein and not type H" (drops waters, ions, hydrogens)
            key = (line[21], int(line[22:26]), line[26])
            if key not in res_keys:
                res_keys[key] = len(res_keys)
            serial = int(line[6:11])
            serial_to_idx[serial] = len(atoms)
            atoms.append(Atom(serial, name, element, res_name, key[1], key[0], res_keys[key]))
            xyz.append([float(line[30:38]) / 10.0, float(line[38:46]) / 10.0, float(line[46:54]) / 10.0])
        elif rec == "CONECT":
            nums = [int(line[k:k + 5]) for k in range(6, len(line.rstrip()), 5) if line[k:k + 5].strip()]
            conect += [(nums[0], b) for b in nums[1:]]
    top = Topology(atoms=atoms)
    by_res: Dict[int, Dict[str, int]] = {}
    for i, a in enumerate(atoms):
        by_res.setdefault(a.residue_index, {})[a.name] = i
    seen = set()

    def add(i: int, j: int) -> None:
        if (i, j) not in seen and (j, i) not in seen:
            seen.add((i, j))
            top.bonds.append((i, j))

    res_order = sorted(by_res)
    for r in res_order:
        names = by_res[r]
        res_name = atoms[next(iter(names.values()))].res_name
        pairs = _SIDE[res_name] if res_name in ("ACE", "NME") else _BB + _SIDE[res_name]
        for a, b in pairs:
            if a in names and b in names:
                add(names[a], names[b])
    for r0, r1 in zip(res_order[:-1], res_order[1:]):  # peptide bonds within a chain
        a0, a1 = by_res[r0], by_res[r1]
        if "C" in a0 and "N" in a1 and atoms[a0["C"]].chain == atoms[a1["N"]].chain:
            add(a0["C"], a1["N"])
    for s0, s1 in conect:
        if s0 in serial_to_idx and s1 in serial_to_idx:
            add(serial_to_idx[s0], serial_to_idx[s1])
    return top, np.asarray(xyz, dtype=np.float32).reshape(-1, 3)


def preprocess_topology(top: Topology, positions: Optional[np.ndarray] = None, label: str = "", loss_weight: float = 1.0
                        ) -> DataWithResidueInformation:
    """Graph tensors of one peptide (data/_mdtraj.py:56-89): index vocabularies of ResidueMetadata, bonded edge_index."""
    g = DataWithResidueInformation(
        atom_type_index=torch.tensor([encode_atom_type(a.element) for a in top.atoms], dtype=torch.int32),
        residue_code_index=torch.tensor([encode_residue(a.res_name) for a in top.atoms], dtype=torch.int32),
        residue_sequence_index=torch.tensor([a.residue_index for a in top.atoms], dtype=torch.int32),
        atom_code_index=torch.tensor([encode_atom_code(a.name) for a in top.atoms], dtype=torch.int32),
        edge_index=torch.tensor(top.bonds, dtype=torch.long).reshape(-1, 2).T.contiguous(),
        pos=None if positions is None else torch.as_tensor(positions, dtype=torch.float32))
    g["residue_index"] = g["residue_sequence_index"]
    g["num_residues"] = int(g["residue_sequence_index"].max().item()) + 1 if top.n_atoms else 0
    g["residues"] = [a.res_name for a in top.atoms]
    g["atom_names"] = [a.name for a in top.atoms]
    g["dataset_label"] = label
    g["loss_weight"] = torch.tensor([loss_weight], dtype=torch.float32)
    return g


def graph_from_pdb(path_or_text: str, label: str = "", loss_weight: float = 1.0) -> Tuple[DataWithResidueInformation, Topology]:
    top, xyz = read_pdb(path_or_text)
    return preprocess_topology(top, xyz, label=label, loss_weight=loss_weight), top


# ---- writers ---------------------------------------------------------------------------------------------------------
def write_pdb(path: str, top: Topology, frames_nm: np.ndarray) -> None:
    """Multi-MODEL PDB of frames [T, n, 3] (nm in, Angstrom out)."""
    frames = np.asarray(frames_nm, dtype=np.float64).reshape(-1, top.n_atoms, 3) * 10.0
    with open(path, "w") as f:
        for m, xyz in enumerate(frames, start=1):
            f.write(f"MODEL     {m:4d}\n")
            for i, (a, p) in enumerate(zip(top.atoms, xyz), start=1):
                name = a.name if len(a.name) == 4 else f" {a.name:<3s}"
                f.write(f"ATOM  {i:5d} {name} {a.res_name:>3s} {a.chain or 'A'}{a.res_seq:4d}    {p[0]:8.3f}{p[1]:8.3f}{p[2]:8.3f}"
                        f"  1.00  0.00          {a.element:>2s}\n")
            f.write("ENDMDL\n")
        for i, j in top.bonds:
            f.write(f"CONECT{i + 1:5d}{j + 1:5d}\n")
        f.write("END\n")


def write_dcd(path: str, frames_nm: np.ndarray) -> None:
    """CHARMM-format DCD (little endian, no unit cell) of frames [T, n, 3] (nm in, Angstrom out)."""
    frames = np.asarray(frames_nm, dtype=np.float32) * 10.0
    T, n = frames.shape[0], frames.shape[1]

    def block(payload: bytes) -> bytes:
        return struct.pack("<i", len(payload)) + payload + struct.pack("<i", len(payload))

    icntrl = [0] * 20
    icntrl[0], icntrl[1], icntrl[2], icntrl[3], icntrl[19] = T, 0, 1, T, 24
    with open(path, "wb") as f:
        f.write(block(b"CORD" + struct.pack("<9i", *icntrl[:9]) + struct.pack("<f", 1.0) + struct.pack("<10i", *icntrl[10:])))
        title = b"Created by jamun_b200".ljust(80)
        f.write(block(struct.pack("<i", 1) + title))
        f.write(block(struct.pack("<i", n)))
        for fr in frames:
            for k in range(3):
                f.write(block(np.ascontiguousarray(fr[:, k]).astype("<f4").tobytes()))


def read_dcd(path: str) -> np.ndarray:
    """Frames [T, n, 3] in nm of a DCD written by write_dcd (round-trip check / analysis)."""
    raw = open(path, "rb").read()
    off = 0

    def block():
        nonlocal off
        (n,) = struct.unpack_from("<i", raw, off)
        payload = raw[off + 4: off + 4 + n]
        off += 8 + n
        return payload

    head = block()
    T = struct.unpack_from("<i", head, 4)[0]
    block()
    n = struct.unpack("<i", block())[0]
    out = np.zeros((T, n, 3), dtype=np.float32)
    for t in range(T):
        for k in range(3):
            out[t, :, k] = np.frombuffer(block(), dtype="<f4")
    return out / 10.0
