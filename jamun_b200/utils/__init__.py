from .align import align_A_to_B_batched, kabsch_algorithm
from .mean_center import mean_center
from .sampling_wrapper import ModelSamplingWrapper
from .unsqueeze_trailing import unsqueeze_trailing
from ..data import Batch, Data, DataWithResidueInformation

__all__ = ["align_A_to_B_batched", "kabsch_algorithm", "mean_center", "ModelSamplingWrapper", "unsqueeze_trailing",
           "Batch", "Data", "DataWithResidueInformation"]
