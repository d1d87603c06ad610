#!/usr/bin/env python
"""bench.py -- walk-jump chain-steps/s (atoms x steps / s) of the B200-native hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (the reference algorithm on the host CPU: oracle port)

A bench "step" is one pass of the hot path over one batch: `--inner` consecutive BAOAB walk-jump steps (radius graph,
edge features, 6 conv blocks, head, fused integrator + jump) over all chains of the workload.  The workload at N=1 is
BASELINE.json configs[1]: uncapped 2AA peptides, 1024 chains (per GPU; weak scaling at N>1), random-init default
denoiser with output_gain=1, synthetic coordinates.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SIGMA = 0.04
MCMC = dict(delta=0.04, friction=1.0, M=1.0, inverse_temperature=1.0, score_fn_clip=100.0)
METRIC = "walkjump_atom_steps_per_s"
UNIT = "atoms*steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="2AA", choices=["2AA", "4AA", "ala2_capped", "protein1000"])
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (default: 1024; protein1000: 64)")
    ap.add_argument("--inner", type=int, default=32, help="walk-jump steps per bench step")
    ap.add_argument("--cpu-sample-chains", type=int, default=24)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def measured_traffic(args, atoms):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture (same workload only)."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        if args.workload == "2AA" and (args.chains or 1024) == 1024:
            return d["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_workload(args, rank: int):
    from jamun_b200 import synthetic

    chains = args.chains or (64 if args.workload == "protein1000" else 1024)
    sizes_all = synthetic.workload_sizes(args.workload, chains * max(1, args.gpus))
    sizes = sizes_all[rank * chains:(rank + 1) * chains]
    n_res = {"2AA": 2, "4AA": 4, "ala2_capped": 2, "protein1000": 100}[args.workload]
    t = synthetic.make_tensors(sizes, n_res=n_res, first_chain_id=rank * chains)
    return t, sizes


def oracle_model():
    from oracle import jamun_oracle as O

    torch.manual_seed(0)
    o = O.Denoiser()
    O.randomize_for_parity(o)
    return o


def cpu_reference_rate(args, sizes_sample, steps: int):
    """The reference's formulation (oracle port) on the host cores: atoms x steps / s on a bounded sample."""
    from jamun_b200 import synthetic
    from oracle import jamun_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    o = oracle_model()
    n_res = {"2AA": 2, "4AA": 4, "ala2_capped": 2, "protein1000": 100}[args.workload]
    t = synthetic.make_tensors(sizes_sample, n_res=n_res)
    ob = O.OracleBatch(pos=t["pos"], batch=t["batch"], num_graphs=t["num_graphs"], edge_index=t["edge_index"],
                       atom_type_index=t["atom_type_index"], atom_code_index=t["atom_code_index"],
                       residue_code_index=t["residue_code_index"], residue_sequence_index=t["residue_sequence_index"])
    y0 = t["pos"] + SIGMA * torch.randn(t["pos"].shape)
    with torch.no_grad():
        O.walk_jump(o, ob, y0, SIGMA, mcmc=O.baoab, v_init="gaussian", steps=2, save_trajectory=True, redundant_jump=False, **MCMC)
        t0 = time.perf_counter()
        # steps+1 score evaluations advance `steps` Langevin updates; the reference's redundant jump pass doubles that
        O.walk_jump(o, ob, y0, SIGMA, mcmc=O.baoab, v_init="gaussian", steps=steps + 1, save_trajectory=True,
                    redundant_jump=True, **MCMC)
        dt = time.perf_counter() - t0
    atoms = int(t["pos"].shape[0])
    return atoms * steps / dt, atoms, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from jamun_b200 import synthetic

    chains = args.chains or (64 if args.workload == "protein1000" else 1024)
    sizes = synthetic.workload_sizes(args.workload, chains)[: args.cpu_sample_chains]
    inner = 1
    rates, times = [], []
    for k in range(args.warmup + args.steps):
        rate, atoms, dt = cpu_reference_rate(args, sizes, inner)
        if k >= args.warmup:
            rates.append(rate)
            times.append(dt)
    value = sum(rates) / len(rates)
    sample = f"first {len(sizes)} chains ({atoms} atoms) of the {args.workload} workload x {inner} walk-jump step(s) per bench step, " \
             f"reference formulation incl. its redundant jump pass"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload} uncapped peptides, {chains} chains/GPU, sigma=0.04, BAOAB", "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_native(args):
    import torch.distributed as dist

    import jamun_b200 as J
    from jamun_b200 import data, ops, utils
    from jamun_b200.sampling.mcmc import BAOAB
    from jamun_b200.sampling.mcmc.functional import fused_baoab
    from jamun_b200.sampling.walkjump import SingleMeasurementSampler

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t, sizes = build_workload(args, rank)
    atoms = int(t["pos"].shape[0])
    # weights: seed-0 default init + parity re-draws (output_gain=1), identical on every rank -- product-side init
    torch.manual_seed(0)
    model = J.default_denoiser()
    gen = torch.Generator().manual_seed(0)
    with torch.no_grad():
        model.arch_module.output_gain.fill_(1.0)
        from jamun_b200.model.noise_conditioning import NoiseConditionalScaling

        for m in model.modules():
            if isinstance(m, NoiseConditionalScaling):
                last = m.scale_predictor[-1]
                last.weight.copy_(torch.randn(last.weight.shape, generator=gen) * 0.1)
                last.bias.copy_(1.0 + torch.randn(last.bias.shape, generator=gen) * 0.1)
    model = model.to(dev).eval()
    torch.manual_seed(1234 + rank)  # per-rank Philox stream (cmdline/sample.py:86-88)
    batch = data.Batch.from_tensors(t).to(dev)
    wrapped = utils.ModelSamplingWrapper(model, batch, SIGMA)
    topo = wrapped.topology
    sampler = SingleMeasurementSampler(BAOAB(steps=args.inner + 1, save_trajectory=False, **MCMC), SIGMA)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    y = wrapped.sample_initial_noisy_positions()

    def device_step():
        # every bench step walks `inner` steps from the same noisy start (same workload per step, as in the e2e leg below;
        # with random-init weights a continued chain drifts apart and its graphs get sparser, i.e. cheaper)
        return fused_baoab(model, topo, y, SIGMA, steps=args.inner + 1, v_init="gaussian", **MCMC)

    for _ in range(args.warmup):
        device_step()
    barrier()
    clocks = ClockSampler(local_rank)
    launches0 = ops.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = device_step()
    if world > 1:  # the run's only collective: final NCCL gather of the denoised samples
        from jamun_b200.sampling import Sampler

        gathered = Sampler.gather_samples_ragged(out["xhat"])
    ev1.record()
    barrier()
    launches = ops.LAUNCHES - launches0
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    tt = torch.tensor([ms, float(atoms)], device=dev, dtype=torch.float64)
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, total_atoms = float(tmax[0]), float(tsum[1])
    else:
        total_atoms = float(atoms)
    value = total_atoms * args.inner * args.steps / (ms * 1e-3)

    # ---------------- end-to-end through the public sampler API with HOST buffers (`e2e`)
    y_host = y.detach().cpu().pin_memory()
    x_host = torch.empty_like(y_host).pin_memory()

    def e2e_step():
        y_dev = y_host.to(dev, non_blocking=True)
        o = sampler.sample(wrapped, y_init=y_dev, v_init="gaussian")
        x_host.copy_(o["sample"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(1, args.warmup)):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ems = e0.elapsed_time(e1)
    if world > 1:
        te = torch.tensor([ems], device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ems = float(te[0])
    e2e_value = total_atoms * args.inner * args.steps / (ems * 1e-3)
    bytes_io = int(y_host.numel() * 4)

    # ---------------- roofline of the dominant kernels, each timed alone with CUDA events on the launch stream
    pk, pk_src = peaks()
    plan = model.arch_module.plan(model.sigma_context(SIGMA).c_noise, dev)
    blk = plan.blocks[1]
    x_in = topo.xs[0]
    E = int(topo.rowptr[-1].item())
    deg = E / max(1, atoms)

    def time_kernel(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / reps

    from jamun_b200 import engine as _engine

    rp = topo.chunk_rows
    st0, st1 = 65 * 5, 65 * 2
    a1_off, comp = st0 * rp * 32, st1 * rp * 32
    _engine.conv_tc(topo, blk, x_in, topo.conv)
    nrows = min(rp, atoms)
    base = topo.a_ws.data_ptr()
    _engine.conv_tc_join(topo, blk)
    build_ms = time_kernel(lambda: ops.conv_build_tc(x_in, 120, 32, topo.rowptr, topo.col, topo.h, topo.rhat, 0, nrows, rp, base,
                                                     base + 4 * a1_off, comp, topo.inv_deg))
    p2_ms = time_kernel(lambda: ops.conv_p2(topo.rowptr, topo.src_rowptr, topo.src_eid, topo.h, topo.rhat, topo.y, topo.t_edge,
                                            topo.p2.data_ptr(), 96, blk["alpha1"]))
    gemm_ms = time_kernel(lambda: ops.gemm_tf32x3(
        [base] + [base + 4 * (a1_off + c * comp) for c in range(3)], [blk["b0_img"].data_ptr()] + [blk["b1_img"].data_ptr()] * 3,
        [st0, st1, st1, st1], [160, 32, 32, 32], [152, 32, 32, 32], [0, 152, 184, 216], [1.0] * 4, nrows, rp,
        topo.inv_deg.data_ptr(), topo.conv.data_ptr(), 248))
    rows_all = (atoms + 127) // 128 * 128
    ygemm_ms = time_kernel(lambda: ops.gemm_tf32x3([topo.xs_op.data_ptr()], [blk["wy_img"].data_ptr()], [4], [128], [128], [0], [1.0],
                                                   atoms, rows_all, None, topo.y.data_ptr(), _engine.Y_LD, col_blocks=17,
                                                   b_block_floats=4 * 2 * 128 * 32))
    # algorithmic work per launch: the aggregated contraction 2*65*(152*152 + 3*64*32) FLOP per atom on the
    # tensor pipe (the x3 TF32 passes needed for fp32 parity are NOT counted); its A operand 65*(160+3*64)*4 B per atom via HBM
    gemm_flop = nrows * 2.0 * 65 * (152 * 152 + 3 * 64 * 32)
    a_bytes = nrows * 65.0 * (160 + 3 * 64) * 4
    tf32_peak = pk["bf16_tflops"] / 2.0
    roofline = {"kernel": "gemm_tf32x3_kernel (hidden ConvBlock contraction, tcgen05 kind::tf32, 3xTF32)", "bound": "tensor",
                "achieved": gemm_flop / (gemm_ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": gemm_flop / (gemm_ms * 1e-3) / 1e12 / tf32_peak, "traffic": measured_traffic(args, atoms),
                "peak_source": f"{pk_src} bf16 burst / 2 (dense tf32 rate; fp32-parity 3xTF32 needs 3 passes, so 1/3 is the ceiling)",
                "ms_per_launch": gemm_ms, "hbm_GBps_A_operand": a_bytes / (gemm_ms * 1e-3) / 1e9,
                "hbm_frac_of_measured": a_bytes / (gemm_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "mean_in_degree": deg,
                "second_kernel": {"kernel": "conv_build_tc_kernel<120,32> (per-node aggregate F^T.H on tcgen05, 3xTF32; writes the A operand)",
                                  "bound": "hbm", "ms_per_launch": build_ms, "achieved": a_bytes / (build_ms * 1e-3) / 1e9,
                                  "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": a_bytes / (build_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                                  "tensor_TFLOPs": nrows * deg * 2 * 65 * (152 + 3 * 64) / (build_ms * 1e-3) / 1e12},
                "fourth_kernel": {"kernel": "conv_p2_edge_kernel + conv_p2_reduce_kernel (0e(x)1e->1e, source-major)",
                                  "ms_per_launch": p2_ms, "fma_TFLOPs": nrows * deg * 2 * 65 * 32 / (p2_ms * 1e-3) / 1e12},
                "third_kernel": {"kernel": "gemm_tf32x3_kernel, 17 column-block passes, A-stationary (per-node transform Y = x_s.W, N=2080)",
                                 "ms_per_launch": ygemm_ms, "achieved": atoms * 2.0 * 120 * 2080 / (ygemm_ms * 1e-3) / 1e12,
                                 "unit": "TFLOP/s"}}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{args.workload} uncapped peptides, {len(sizes)} chains/GPU ({atoms} atoms/GPU), "
                                       f"{args.inner} BAOAB walk-jump steps per bench step, sigma=0.04, default e3conv denoiser "
                                       f"(random init, output_gain=1)",
                           "l2": "inputs larger than L2: every denoiser evaluation streams a 1.7 GB conv operand per layer "
                                 "(>> 126 MB L2) and every step advances y, so nothing is served from a warm cache",
                           "parallelism": f"chains sharded x{world}, final NCCL all_gather of samples"},
                "clocks": clk, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_io, "d2h_bytes_per_step": bytes_io,
                        "ms_per_step": ems / args.steps},
                "roofline": roofline}
        if not args.no_cpu_baseline and world == 1:
            from jamun_b200 import synthetic

            chains = args.chains or (64 if args.workload == "protein1000" else 1024)
            sample_sizes = synthetic.workload_sizes(args.workload, chains)[: args.cpu_sample_chains]
            rate, a, dt = cpu_reference_rate(args, sample_sizes, 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"first {len(sample_sizes)} chains ({a} atoms) x 1 walk-jump step, reference "
                                              f"formulation incl. redundant jump pass, {dt:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
