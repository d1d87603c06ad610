"""Vocabulary of the atom / residue indices the denoiser embeds (mirror of /root/reference/src/jamun/utils/residue_metadata.py:
the index of a symbol in these lists *is* the embedding row a released checkpoint learned, so the lists are part of the
checkpoint format)."""
from __future__ import annotations

from typing import Dict, List


class ResidueMetadata:
    ATOM_TYPES: List[str] = ["C", "O", "N", "F", "S"]
    ATOM_CODES: List[str] = ["C", "O", "N", "S", "CA", "CB"]
    RESIDUE_CODES: List[str] = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLU", "GLN", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE",
                                "PRO", "SER", "THR", "TRP", "TYR", "VAL", "ACE", "NME"]
    AA_3CODES: Dict[str, str] = dict(zip("ARNDCEQGHILKMFPSTWYV", RESIDUE_CODES[:20]))
    AA_1CODES: Dict[str, str] = {v: k for k, v in AA_3CODES.items()}


def _index_or_len(table: List[str], key: str) -> int:
    return table.index(key) if key in table else len(table)  # unknown symbols share one extra row


def encode_atom_type(atom_type: str) -> int:
    return _index_or_len(ResidueMetadata.ATOM_TYPES, atom_type)


def encode_atom_code(atom_code: str) -> int:
    return _index_or_len(ResidueMetadata.ATOM_CODES, atom_code)


def encode_residue(residue_name: str) -> int:
    return _index_or_len(ResidueMetadata.RESIDUE_CODES, residue_name)


def convert_to_three_letter_code(aa: str) -> str:
    aa = aa.upper()
    if len(aa) == 1 and aa in ResidueMetadata.AA_3CODES:
        return ResidueMetadata.AA_3CODES[aa]
    if len(aa) == 3 and aa in ResidueMetadata.AA_1CODES:
        return aa
    raise ValueError(f"Invalid amino acid code: {aa}")


def convert_to_three_letter_codes(peptide: str) -> str:
    return peptide if "_" in peptide else "_".join(convert_to_three_letter_code(aa) for aa in peptide)


def convert_to_one_letter_code(aa: str) -> str:
    aa = aa.upper()
    if len(aa) == 1 and aa in ResidueMetadata.AA_3CODES:
        return aa
    if len(aa) == 3 and aa in ResidueMetadata.AA_1CODES:
        return ResidueMetadata.AA_1CODES[aa]
    raise ValueError(f"Invalid amino acid code: {aa}")


def convert_to_one_letter_codes(peptide: str) -> str:
    return peptide if "_" not in peptide else "".join(convert_to_one_letter_code(aa) for aa in peptide.split("_"))
