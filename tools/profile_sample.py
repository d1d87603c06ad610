"""Per-kernel device time of one denoiser evaluation + walk step (torch.profiler, CUDA activity, eager launches).
    python tools/profile_sample.py [workload] [chains]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from jamun_b200 import data, utils  # noqa: E402
from jamun_b200.sampling.mcmc.functional import fused_baoab  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "2AA"
chains = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEFAULT_CHAINS[workload]
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
t, sizes = bench.workload_tensors(workload, chains, 0, 1)
batch = data.Batch.from_tensors(t).to(dev)
wrapped = utils.ModelSamplingWrapper(model, batch, bench.SIGMA)
y = wrapped.sample_initial_noisy_positions()
run = lambda: fused_baoab(model, wrapped.topology, y, bench.SIGMA, steps=3, v_init="gaussian", use_cuda_graph=False, **bench.MCMC)  # noqa: E731
run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run()
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        c = rows.setdefault(e.name[:100], [0, 0.0])
        c[0] += 1
        c[1] += e.device_time
total = sum(v[1] for v in rows.values())
print(f"{workload} x {chains} chains, {t['pos'].shape[0]} atoms: {total / 3e3:.3f} ms of kernels per step (3 eager steps)")
for k, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{us / 3e3:9.4f} ms/step {100 * us / total:5.1f}%  x{n:<4d} {k}")
