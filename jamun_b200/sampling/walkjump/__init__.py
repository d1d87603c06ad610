from ._single_measurement import SingleMeasurementSampler

__all__ = ["SingleMeasurementSampler"]
