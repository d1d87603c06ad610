"""`jamun_sample` -- walk-jump sampling from initial structures (mirror of /root/reference/src/jamun/cmdline/sample.py:41-124
without Hydra: the same steps -- model from a checkpoint, initial graphs from the init datasets, callbacks, Sampler,
per-rank seed, optional fine-tuning on the initial structures, `sampler.sample(...)` -- driven by argparse flags or by
Hydra-style `key=value` overrides of the reference's config keys (`num_batches=4 continue_chain=false seed=1`).

    python -m jamun_b200.cmdline.sample --pdb init.pdb --checkpoint model.ckpt --num-batches 10 --steps 1000 --output-dir run/
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import List, Optional, Sequence

import torch

_BOOL = {"true": True, "false": False, "1": True, "0": False, "yes": True, "no": False}
_OVERRIDES = {  # reference config key -> argparse destination
    "num_batches": "num_batches", "continue_chain": "continue_chain", "repeat_init_samples": "repeat_init_samples", "seed": "seed",
    "sigma": "sigma", "delta": "delta", "friction": "friction", "M": "M", "inverse_temperature": "inverse_temperature",
    "score_fn_clip": "score_fn_clip", "num_sampling_steps_per_batch": "steps", "save_every_n_steps": "save_every_n_steps",
    "burn_in_steps": "burn_in_steps", "checkpoint_dir": "checkpoint", "sample_pdb": "pdb", "finetune_on_init.num_steps": "finetune_steps",
}


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="jamun_sample", description=__doc__.split("\n\n")[0])
    ap.add_argument("--pdb", nargs="+", default=None, help="initial structure(s); heavy protein atoms are used")
    ap.add_argument("--checkpoint", default=None, help="Lightning .ckpt of the denoiser (omit with --random-init)")
    ap.add_argument("--random-init", action="store_true", help="default architecture with random weights (smoke tests, benchmarks)")
    ap.add_argument("--num-batches", type=int, default=1)
    ap.add_argument("--repeat-init-samples", type=int, default=1)
    ap.add_argument("--continue-chain", type=lambda s: _BOOL[s.lower()], default=True)
    ap.add_argument("--sigma", type=float, default=0.04)
    ap.add_argument("--delta", type=float, default=0.04)
    ap.add_argument("--friction", type=float, default=1.0)
    ap.add_argument("--M", type=float, default=1.0)
    ap.add_argument("--inverse-temperature", type=float, default=1.0)
    ap.add_argument("--score-fn-clip", type=float, default=100.0)
    ap.add_argument("--mcmc", choices=["baoab", "aboba"], default="baoab")
    ap.add_argument("--steps", type=int, default=1000, help="Langevin steps per batch (num_sampling_steps_per_batch)")
    ap.add_argument("--save-every-n-steps", type=int, default=1)
    ap.add_argument("--burn-in-steps", type=int, default=0)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--output-dir", default=".")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--finetune-steps", type=int, default=0, help="test-time adaptation steps on the initial structures")
    ap.add_argument("--finetune-lr", type=float, default=None)
    ap.add_argument("overrides", nargs="*", help="Hydra-style key=value overrides of the reference's config keys")
    return ap


def parse_args(argv: Optional[Sequence[str]] = None) -> argparse.Namespace:
    ap = build_parser()
    args = ap.parse_args(argv)
    for tok in args.overrides:
        if "=" not in tok:
            ap.error(f"unrecognised argument {tok!r} (expected key=value)")
        key, val = tok.split("=", 1)
        key = key.lstrip("+").split("batch_sampler.mcmc.")[-1].split("batch_sampler.")[-1]
        if key not in _OVERRIDES:
            ap.error(f"unknown config key {key!r}; known: {sorted(_OVERRIDES)}")
        dest = _OVERRIDES[key]
        cur = getattr(args, dest)
        if dest == "pdb":
            setattr(args, dest, [val])
        elif isinstance(cur, bool):
            setattr(args, dest, _BOOL[val.lower()])
        elif isinstance(cur, int) and not isinstance(cur, bool):
            setattr(args, dest, int(val))
        elif isinstance(cur, float):
            setattr(args, dest, float(val))
        else:
            setattr(args, dest, val)
    if not args.pdb:
        ap.error("at least one --pdb (sample_pdb=...) initial structure is required")
    if not args.checkpoint and not args.random_init:
        ap.error("give --checkpoint (checkpoint_dir=...) or --random-init")
    return args


def get_initial_graphs(graphs: List, repeat: int = 1):
    """cmdline/sample.py:27-38: every initial graph `repeat` times, batched."""
    from ..data import Batch

    return Batch.from_data_list([g for g in graphs for _ in range(repeat)])


def find_checkpoint(path: str) -> str:
    """utils/checkpoint.py: a .ckpt file, or a directory holding checkpoints (the most recent `last.ckpt` / newest file)."""
    if os.path.isdir(path):
        cands = sorted((os.path.join(r, f) for r, _, fs in os.walk(path) for f in fs if f.endswith(".ckpt")), key=os.path.getmtime)
        if not cands:
            raise FileNotFoundError(f"no .ckpt under {path}")
        last = [c for c in cands if os.path.basename(c) == "last.ckpt"]
        return (last or cands)[-1]
    return path


def run(args: argparse.Namespace):
    import jamun_b200 as J
    from ..callbacks import SaveTrajectoryCallback
    from ..metrics import SaveTrajectory
    from ..pdb import graph_from_pdb
    from ..sampling import Sampler
    from ..sampling.mcmc import ABOBA, BAOAB
    from ..sampling.walkjump import SingleMeasurementSampler

    if args.checkpoint:
        model = J.model.Denoiser.load_from_checkpoint(find_checkpoint(args.checkpoint))
    else:
        torch.manual_seed(0)
        model = J.default_denoiser(sigma=args.sigma)
    graphs, metrics = [], {}
    for path in args.pdb:
        label = os.path.splitext(os.path.basename(path))[0]
        graph, top = graph_from_pdb(path, label=label)
        graphs.append(graph)
        metrics[label] = SaveTrajectory(label, top, output_root=args.output_dir, init_positions_nm=graph.pos.numpy())
    init_graphs = get_initial_graphs(graphs, repeat=args.repeat_init_samples)
    callbacks = [SaveTrajectoryCallback(metrics)]
    sampler = Sampler(devices=1, precision="32-true", callbacks=callbacks)
    mcmc_cls = BAOAB if args.mcmc == "baoab" else ABOBA
    mcmc = mcmc_cls(delta=args.delta, friction=args.friction, M=args.M, steps=args.steps, save_trajectory=True,
                    save_every_n_steps=args.save_every_n_steps, burn_in_steps=args.burn_in_steps,
                    inverse_temperature=args.inverse_temperature, score_fn_clip=args.score_fn_clip)
    batch_sampler = SingleMeasurementSampler(mcmc, args.sigma)
    # during sampling, ranks generate different chains (cmdline/sample.py:86-88)
    torch.manual_seed(args.seed + sampler.fabric.global_rank)
    if args.finetune_steps:  # test-time adaptation on the initial structures (cmdline/sample.py:91-116)
        model = model.to(sampler.fabric.device).train()
        before = float(sum(p.detach().double().sum() for p in model.parameters()))
        optim = model.configure_optimizers()["optimizer"]
        if args.finetune_lr:
            for grp in optim.param_groups:
                grp["lr"] = args.finetune_lr
        batch = init_graphs.to(sampler.fabric.device)
        for it in range(args.finetune_steps):
            optim.zero_grad()
            model.training_step(batch, it)["loss"].backward()
            optim.step()
        after = float(sum(p.detach().double().sum() for p in model.parameters()))
        print(f"Model parameters changed: {before} -> {after}", file=sys.stderr)
        model.eval()
    sampler.sample(model=model, batch_sampler=batch_sampler, init_graphs=init_graphs, num_batches=args.num_batches,
                   continue_chain=args.continue_chain)
    return metrics


def main(argv: Optional[Sequence[str]] = None) -> int:
    run(parse_args(argv))
    return 0


if __name__ == "__main__":
    sys.exit(main())
