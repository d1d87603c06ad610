#!/usr/bin/env python
"""Per-kernel-family device time of one denoiser evaluation + walk step, measured with CUDA events around every C-ABI call
(eager launches, warm caches).  Usage: python tools/profile_eval.py --workload 2AA|4AA|protein1000 [--chains N]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import jamun_b200 as J  # noqa: E402
from jamun_b200 import data, ops, synthetic, utils  # noqa: E402
from jamun_b200.sampling.mcmc.functional import fused_baoab  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="2AA")
ap.add_argument("--chains", type=int, default=None)
args = ap.parse_args()
chains = args.chains or (64 if args.workload == "protein1000" else 1024)
sizes = synthetic.workload_sizes(args.workload, chains)
t = synthetic.make_tensors(sizes, n_res={"2AA": 2, "4AA": 4, "protein1000": 100}.get(args.workload, 2))
torch.manual_seed(0)
model = J.default_denoiser()
with torch.no_grad():
    model.arch_module.output_gain.fill_(1.0)
model = model.cuda().eval()
batch = data.Batch.from_tensors(t).to("cuda")
wrapped = utils.ModelSamplingWrapper(model, batch, 0.04)
y = wrapped.sample_initial_noisy_positions()
kw = dict(delta=0.04, friction=1.0, M=1.0, inverse_temperature=1.0, score_fn_clip=100.0)
fused_baoab(model, wrapped.topology, y, 0.04, steps=3, v_init="gaussian", use_cuda_graph=False, **kw)  # warm-up

records = collections.defaultdict(list)
names = ["center_scale", "radius_csr", "csr_by_source", "edge_geom", "edge_radial_hidden", "edge_radial_hidden_all", "pack_rows", "gemm_tf32x3", "gemm_f16x3", "gemm_f16x3_fused", "conv_build_a", "conv_build_tc", "conv_p2", "block_tail", "tail_pack", "tail_mix",
         "head", "walk_step", "gaussian_axpy"]
orig = {n: getattr(ops, n) for n in names}


def wrap(n):
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig[n](*a, **k)
        e1.record()
        tag = n
        if n == "gemm_f16x3_fused":
            tag = "gemm_f16x3_fused (contraction + gate)" if a[10]["mode"] == 1 else "gemm_f16x3_fused (block tail + mix)"
        if n in ("gemm_tf32x3", "gemm_f16x3"):
            tag = (f"{n} (transform Y)" if k.get("col_blocks", 1) > 1 else
                   f"{n} (block tail)" if a[3][0] == 128 else f"{n} (contraction)")
        records[tag].append((e0, e1))
        return r
    return f


for n in names:
    setattr(ops, n, wrap(n))
steps = 4
fused_baoab(model, wrapped.topology, y, 0.04, steps=steps, v_init="gaussian", use_cuda_graph=False, **kw)
torch.cuda.synchronize()
tot = 0.0
rows = []
for tag, evs in records.items():
    ms = sum(a.elapsed_time(b) for a, b in evs)
    rows.append((ms / steps, len(evs) / steps, tag))
    tot += ms / steps
print(f"workload {args.workload}: {t['pos'].shape[0]} atoms, {int(wrapped.topology.rowptr[-1])} edges; per evaluation+step {tot:.3f} ms "
      f"(sum of kernel families, eager)")
for ms, n, tag in sorted(rows, reverse=True):
    print(f"  {tag:32s} {ms:8.3f} ms  {100 * ms / tot:5.1f}%   ({n:.1f} launches)")
