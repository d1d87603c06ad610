"""mean_center (mirror of /root/reference/src/jamun/utils/mean_center.py:7-12) on the K0 kernel."""
from __future__ import annotations

from .. import engine, ops


def mean_center(x):
    """Per-graph centroid subtraction; returns a shallow clone with new `pos`."""
    out = x.clone("pos")
    if "_topology" not in x:
        x["_topology"] = engine.Topology(x, x.pos.device)  # cached on the input too: later calls on the same batch reuse it
    topo = out["_topology"] = x["_topology"]
    ybar, _ = ops.center_scale(x.pos.contiguous(), topo.chain_ptr, 1.0)
    out.pos = ybar
    return out
