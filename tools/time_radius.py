"""Brute-force scan vs cell-list search (jamun_radius_csr vs jamun_radius_csr_cells) on equal-length chains: ms per CSR build.
    python tools/time_radius.py 1000 512   (chain length, chains)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from jamun_b200 import data, engine, synthetic  # noqa: E402

n, chains = int(sys.argv[1]), int(sys.argv[2])
one = synthetic.make_tensors([n] * 2)
t = synthetic.make_tensors([n] * 2)
# replicate two generated chains (chain generation is O(n^2) on the host)
reps = chains // 2
pos = one["pos"].repeat(reps, 1)
batch = torch.arange(2 * reps).repeat_interleave(n)
ei = torch.cat([one["edge_index"] + 2 * n * r for r in range(reps)], dim=1)
tt = dict(pos=pos, batch=batch, edge_index=ei, num_graphs=2 * reps, ptr=torch.arange(2 * reps + 1) * n,
          loss_weight=torch.ones(2 * reps))
for k in ("atom_type_index", "atom_code_index", "residue_code_index", "residue_sequence_index"):
    tt[k] = one[k].repeat(reps)
res = {}
for impl in ("brute", "cells"):
    os.environ["JAMUN_B200_RADIUS"] = impl
    topo = engine.Topology(data.Batch.from_tensors(tt), "cuda")
    p = pos.cuda()
    for _ in range(3):
        topo.build_csr(p, 0.5872643)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        topo.build_csr(p, 0.5872643)
    e1.record()
    torch.cuda.synchronize()
    res[impl] = e0.elapsed_time(e1) / 10
print(f"chain length {n} x {2 * reps} chains ({2 * reps * n} atoms): brute {res['brute']:.3f} ms, cells {res['cells']:.3f} ms (incl. csr_by_source)")
