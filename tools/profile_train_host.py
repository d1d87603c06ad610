"""Host-side profile (cProfile) of training steps at the reference's batch size (32 graphs): python tools/profile_train_host.py"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from jamun_b200 import data  # noqa: E402

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
model = bench.make_model(dev).train()
t, sizes = bench.workload_tensors("train4AA", graphs, 0, 1)
batch = data.Batch.from_tensors(t).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-4)


def step():
    opt.zero_grad(set_to_none=True)
    model.training_step(batch, 0)["loss"].backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
