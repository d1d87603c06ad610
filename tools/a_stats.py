import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from jamun_b200 import data, engine, ops, utils
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
for wl in ("2AA", "4AA"):
    t, sizes = bench.workload_tensors(wl, 256, 0, 1)
    batch = data.Batch.from_tensors(t).to(dev)
    wrapped = utils.ModelSamplingWrapper(model, batch, bench.SIGMA)
    topo = wrapped.topology
    y = wrapped.sample_initial_noisy_positions()
    plan = model.arch_module.plan(model.sigma_context(bench.SIGMA).c_noise, dev)
    # hook conv_build_tc to record stats of the operand after each build
    orig = ops.conv_build_tc
    stats = []
    def hooked(*a, **k):
        r = orig(*a, **k)
        torch.cuda.synchronize()
        N = topo.N
        rp = topo.chunk_rows
        s_in = a[1]
        nst = 65 * (11 if s_in == 120 else 2)
        A = topo.a_ws[: nst * rp * 32].abs()
        A = A[torch.isfinite(A)]
        nz = A[A > 0]
        q = torch.quantile(nz[torch.randint(0, nz.numel(), (1_000_000,), device=dev)], torch.tensor([0.01, 0.1, 0.5, 0.9, 0.99], device=dev))
        stats.append((s_in, float(A.max()), [float(v) for v in q]))
        return r
    ops.conv_build_tc = hooked
    wrapped.xhat(y, bench.SIGMA)
    ops.conv_build_tc = orig
    for i, (s_in, mx, q) in enumerate(stats):
        print(wl, "layer", i, "max |A| %.3g" % mx, "quantiles 1/10/50/90/99%%: " + " ".join("%.2e" % v for v in q))
