from .align import align_A_to_B_batched, kabsch_algorithm
from .average_squared_distance import compute_average_squared_distance, compute_average_squared_distance_from_data
from .mean_center import mean_center
from .residue_metadata import (ResidueMetadata, convert_to_one_letter_code, convert_to_one_letter_codes, convert_to_three_letter_code,
                               convert_to_three_letter_codes, encode_atom_code, encode_atom_type, encode_residue)
from .sampling_wrapper import ModelSamplingWrapper
from .unsqueeze_trailing import unsqueeze_trailing
from ..data import Batch, Data, DataWithResidueInformation

__all__ = ["compute_average_squared_distance", "compute_average_squared_distance_from_data", "ResidueMetadata",
           "encode_atom_type", "encode_atom_code", "encode_residue", "convert_to_three_letter_code", "convert_to_three_letter_codes",
           "convert_to_one_letter_code", "convert_to_one_letter_codes", "align_A_to_B_batched", "kabsch_algorithm", "mean_center", "ModelSamplingWrapper", "unsqueeze_trailing",
           "Batch", "Data", "DataWithResidueInformation"]
