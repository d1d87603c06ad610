from ._save_trajectory import SaveTrajectory

__all__ = ["SaveTrajectory"]
