"""Per-kernel device time of one training step (torch.profiler, CUDA activity only): which backward kernels dominate.
    python tools/profile_train.py [graphs_per_gpu]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from jamun_b200 import data  # noqa: E402

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
model = bench.make_model(dev).train()
t, sizes = bench.workload_tensors("train4AA", graphs, 0, 1)
batch = data.Batch.from_tensors(t).to(dev)
for _ in range(2):
    model.zero_grad()
    model.training_step(batch, 0)["loss"].backward()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.zero_grad()
    model.training_step(batch, 0)["loss"].backward()
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name[:90]
        c = rows.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += e.device_time
total = sum(v[1] for v in rows.values())
print(f"{graphs} graphs, {t['pos'].shape[0]} atoms: {total / 1e3:.2f} ms of kernels per fwd+bwd")
for k, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{us / 1e3:9.3f} ms {100 * us / total:5.1f}%  x{n:<4d} {k}")
