from ._ema import EMA
from ._save_trajectory import SaveTrajectoryCallback

__all__ = ["EMA", "SaveTrajectoryCallback"]
