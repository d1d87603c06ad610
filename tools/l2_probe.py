"""How much faster would the aggregate builder and the contraction GEMM run if the A operand stayed in L2?
Times both kernels on a chunk of rows small enough for its operand to be L2-resident (re-run in a loop: after the first pass the
chunk's lines are in L2) and on the full batch (operand streamed through HBM), per row.
    python tools/l2_probe.py [chunk_rows ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from jamun_b200 import data, engine, ops, utils  # noqa: E402
from jamun_b200.sampling.mcmc.functional import fused_baoab  # noqa: E402

dev = torch.device("cuda", 0)
model = bench.make_model(dev)
t, sizes = bench.workload_tensors("2AA", 1024, 0, 1)
batch = data.Batch.from_tensors(t).to(dev)
wrapped = utils.ModelSamplingWrapper(model, batch, bench.SIGMA)
topo = wrapped.topology
y = wrapped.sample_initial_noisy_positions()
fused_baoab(model, topo, y, bench.SIGMA, steps=3, v_init="gaussian", use_cuda_graph=False, **bench.MCMC)
plan = model.arch_module.plan(model.sigma_context(bench.SIGMA).c_noise, dev)
blk = plan.blocks[1]
x_in = topo.xs[0]
N = topo.N
st0, st1 = 65 * 5, 65 * 2


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def probe(rows):
    rp = (rows + 127) // 128 * 128
    ws = torch.empty((st0 + 3 * st1) * rp * 32, device=dev)
    base = ws.data_ptr()
    a1_off, comp = st0 * rp * 32, st1 * rp * 32
    inv = torch.empty(N, device=dev)
    out = torch.empty(N, 248, device=dev)
    tb = timeit(lambda: ops.conv_build_tc(x_in, 120, 32, topo.rowptr, topo.col, topo.h, topo.rhat, 0, rows, rp, base, base + 4 * a1_off,
                                          comp, inv))
    tiles = rp // 128
    ks = max(1, min(32, 148 // tiles))
    part = torch.empty(ks * rows * 248, device=dev)
    a_ptrs = [base] + [base + 4 * (a1_off + c * comp) for c in range(3)]
    b_ptrs = [blk["b0_img"].data_ptr()] + [blk["b1_img"].data_ptr()] * 3
    kind, sc = blk["gemm_kind"], blk["f16_scales"]
    args = (a_ptrs, b_ptrs, [st0, st1, st1, st1], [160, 32, 32, 32], [152, 32, 32, 32], [0, 152, 184, 216],
            [1.0 / sc[0]] + [1.0 / sc[1]] * 3, rows, rp, inv.data_ptr(), out.data_ptr(), 248)
    if kind == "f16":
        tg = timeit(lambda: ops.gemm_f16x3(*args, k_splits=ks, partial=part if ks > 1 else None))
    elif ks > 1:
        tg = timeit(lambda: ops.gemm_tf32x3_splitk(*args, ks, part))
    else:
        tg = timeit(lambda: ops.gemm_tf32x3(*args))
    mb = (st0 + 3 * st1) * rp * 128 / 1e6
    print(f"rows {rows:6d} (operand {mb:7.1f} MB, {tiles} tiles, split-K {ks:2d}, {tiles * ks:3d} CTAs): builder {tb:8.1f} us = {tb / rows * 1e3:6.1f} ns/row,"
          f" contraction {tg:8.1f} us = {tg / rows * 1e3:6.1f} ns/row")


for rows in [int(a) for a in sys.argv[1:]] or [512, 640, 1024, 2048, 4096, N]:
    probe(min(rows, N))
