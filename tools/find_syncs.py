"""Host synchronisations inside one training step (torch.cuda.set_sync_debug_mode): python tools/find_syncs.py"""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from jamun_b200 import data  # noqa: E402

dev = torch.device("cuda", 0)
model = bench.make_model(dev).train()
t, sizes = bench.workload_tensors("train4AA", 256, 0, 1)
batch = data.Batch.from_tensors(t).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
for _ in range(2):
    model.zero_grad()
    model.training_step(batch, 0)["loss"].backward()
    opt.step()
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    model.zero_grad()
    model.training_step(batch, 0)["loss"].backward()
    opt.step()
torch.cuda.set_sync_debug_mode("default")
print(len(w), "synchronising calls in one step")
for x in w:
    print(f"{x.filename}:{x.lineno}: {str(x.message)[:100]}")
