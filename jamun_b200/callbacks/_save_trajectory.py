"""Sampler callback feeding SaveTrajectory metrics (the reference's callbacks/sampler/_save_trajectory.py protocol: one metric per
dataset label, `update` with every unbatched sample graph, hooks forwarded)."""
from __future__ import annotations

from typing import Dict

from ..metrics import SaveTrajectory


class SaveTrajectoryCallback:
    def __init__(self, metrics: Dict[str, SaveTrajectory]):
        self.metrics = metrics

    def on_sample_start(self, sampler=None):
        for m in self.metrics.values():
            m.on_sample_start()

    def on_before_sample_batch(self, sampler=None):
        pass

    def on_after_sample_batch(self, sample, sampler=None):
        for graph in sample:
            label = graph["dataset_label"] if "dataset_label" in graph else None
            metric = self.metrics.get(label) if label in self.metrics else (next(iter(self.metrics.values())) if len(self.metrics) == 1 else None)
            if metric is None:
                raise ValueError(f"no SaveTrajectory metric for dataset label {label!r}")
            metric.update(graph)
        for m in self.metrics.values():
            m.on_after_sample_batch()

    def on_sample_end(self, sampler=None):
        for m in self.metrics.values():
            m.on_sample_end()
