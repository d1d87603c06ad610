"""Golden vectors produced by EXECUTING the reference's own source files (run in the build container; /root/reference
does not exist on the GPU box, so the outputs are committed as tests/golden/reference_exec.npz).

    python tests/golden/make_reference_golden.py            # writes tests/golden/reference_exec.npz

What runs here is the unmodified code under /root/reference/src/jamun, loaded file by file:

    utils/align.py                       kabsch_algorithm, align_A_to_B_batched
    utils/mean_center.py                 mean_center
    utils/unsqueeze_trailing.py
    sampling/mcmc/functional/_splitting.py   baoab, aboba, create_score_fn, initialize_velocity
    sampling/walkjump/_single_measurement.py SingleMeasurementSampler.sample (walk + the redundant jump pass)
    utils/sampling_wrapper.py            ModelSamplingWrapper.score / .xhat / positions_to_graph
    model/denoiser.py                    Denoiser.normalization_factors, loss_weight, effective_radial_cutoff, add_noise,
                                         xhat, xhat_normalized, score, noise_and_denoise, compute_loss,
                                         noise_and_compute_loss
    model/noise_conditioning.py          NoiseConditionalScaling.scale_predictor, NoiseConditionalSkipConnection weights
    model/atom_embedding.py              AtomEmbeddingWithResidueInformation.forward
    utils/residue_metadata.py            vocabulary tables, encode_* and convert_* helpers (the embedding-row order of checkpoints)
    utils/average_squared_distance.py    compute_distance_matrix, compute_average_squared_distance (numpy)
    distributions/_distributions.py      every sigma / measurement distribution (seeded samples)
    lr_schedules/_lr_schedules.py        the three LambdaLR multipliers

The reference's third-party dependencies are absent here (e3nn, torch_geometric, torch_scatter, torch_cluster, lightning),
so the interpreter gets *stand-ins for the containers and primitives only* (listed in `install_stand_ins`): a Data/Batch
attribute bag with `clone(*keys)`, `scatter_mean`, a LightningModule shell, an `Irreps` that can count irreps, and a
brute-force `radius_graph` whose result the toy network below ignores.  None of the e3nn arithmetic (tensor products,
o3.Linear, Gate, spherical harmonics, radial basis) is executed or pinned by this script -- the E3Conv network is
replaced by a closed-form toy `g(y_scaled, c_noise, r_cut)` (tests/toy_arch.py), so that what IS pinned is everything
around the network: normalisation, centring, cut-off, xhat/score tail, clip, BAOAB/ABOBA, the walk-jump shell, noise,
Kabsch, the loss.
"""
from __future__ import annotations

import copy
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src/jamun"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))  # tests/ (toy_arch)


# ----------------------------------------------------------------------------------------------------------------------
# stand-ins for the absent third-party containers / primitives
# ----------------------------------------------------------------------------------------------------------------------
class Data:
    """torch_geometric.data.Data as an attribute bag: attribute <-> key access, `in`, clone(*keys) = shallow copy with the
    named (or all) tensors cloned, which is the documented behaviour the reference relies on (`x.clone("pos")`)."""

    def __init__(self, **kw):
        object.__setattr__(self, "_store", dict(kw))

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, "_store")[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self._store[k] = v

    def __getitem__(self, k):
        return self._store[k]

    def __setitem__(self, k, v):
        self._store[k] = v

    def __contains__(self, k):
        return k in self._store

    def keys(self):
        return list(self._store)

    def clone(self, *keys):
        out = copy.copy(self)
        object.__setattr__(out, "_store", dict(self._store))
        for k in (keys or self.keys()):
            if torch.is_tensor(out._store[k]):
                out._store[k] = out._store[k].clone()
        return out

    def to(self, *a, **k):
        return self

    @property
    def num_nodes(self):
        return self._store["pos"].shape[0]


class Batch(Data):
    @property
    def num_graphs(self):
        return int(self._store["batch"].max()) + 1


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    dim = dim % src.ndim
    shape = list(src.shape)
    shape[dim] = int(dim_size if dim_size is not None else int(index.max()) + 1)
    res = torch.zeros(shape, dtype=src.dtype)
    return res.index_add_(dim, index, src)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    """torch_scatter.scatter_mean: sum / max(count, 1)."""
    s = scatter_sum(src, index, dim, None, dim_size)
    dim = dim % src.ndim
    cnt = torch.zeros(s.shape[dim], dtype=src.dtype).index_add_(0, index, torch.ones(index.shape[0], dtype=src.dtype))
    view = [1] * s.ndim
    view[dim] = -1
    return s / cnt.clamp(min=1).view(view)


def radius_graph(pos, r, batch, *a, **k):
    """Only so that Denoiser.add_edges runs; the toy network ignores the edges (neighbour lists are NOT pinned here)."""
    d = torch.cdist(pos, pos)
    same = batch[:, None] == batch[None, :]
    m = (d <= r) & same & ~torch.eye(pos.shape[0], dtype=torch.bool)
    dst, src = m.nonzero(as_tuple=True)
    return torch.stack([src, dst])


class Irreps:
    def __init__(self, s):
        self.s = str(s)
        self.num_irreps = sum(int(t.split("x")[0]) for t in self.s.replace(" ", "").split("+"))

    def __repr__(self):
        return self.s


class _NotExecuted(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        raise RuntimeError("e3nn arithmetic is not available here and must not be pinned by this script")


class LightningModule(torch.nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    @property
    def device(self):
        return torch.device("cpu")


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stand_ins():
    tg_data = _module("torch_geometric.data", Data=Data, Batch=Batch)
    tg_nn = _module("torch_geometric.nn", radius_graph=radius_graph)
    _module("torch_geometric", data=tg_data, nn=tg_nn)
    _module("torch_scatter", scatter_mean=scatter_mean, scatter_sum=scatter_sum)
    o3 = _module("e3nn.o3", Irreps=Irreps, ElementwiseTensorProduct=_NotExecuted)
    _module("e3nn", o3=o3)
    plu = _module("lightning.pytorch.utilities", rank_zero_only=lambda f: f)
    plm = _module("lightning.pytorch", LightningModule=LightningModule, LightningDataModule=object, Trainer=object, utilities=plu)
    _module("lightning", pytorch=plm)
    _module("hydra")


def load_reference():
    """Loads the reference files one by one into a synthetic `jamun` package (its own __init__ files import the whole
    application: wandb, mdtraj, hydra, ...)."""
    install_stand_ins()
    pkg = _module("jamun")
    pkg.__path__ = []
    utils = _module("jamun.utils")
    pkg.utils = utils

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m

    ut = load("jamun.utils.unsqueeze_trailing", "utils/unsqueeze_trailing.py")
    mc = load("jamun.utils.mean_center", "utils/mean_center.py")
    al = load("jamun.utils.align", "utils/align.py")
    dw = load("jamun.utils.data_with_residue_info", "utils/data_with_residue_info.py")
    sw = load("jamun.utils.sampling_wrapper", "utils/sampling_wrapper.py")
    utils.unsqueeze_trailing = ut.unsqueeze_trailing
    utils.mean_center = mc.mean_center
    utils.align_A_to_B_batched = al.align_A_to_B_batched
    utils.DataWithResidueInformation = dw.DataWithResidueInformation
    utils.ModelSamplingWrapper = sw.ModelSamplingWrapper
    ref = types.SimpleNamespace(
        align=al, mean_center=mc, sampling_wrapper=sw,
        splitting=load("jamun.sampling.mcmc.functional._splitting", "sampling/mcmc/functional/_splitting.py"),
        single=load("jamun.sampling.walkjump._single_measurement", "sampling/walkjump/_single_measurement.py"),
        denoiser=load("jamun.model.denoiser", "model/denoiser.py"),
        noise=load("jamun.model.noise_conditioning", "model/noise_conditioning.py"),
        embed=load("jamun.model.atom_embedding", "model/atom_embedding.py"),
        meta=load("jamun.utils.residue_metadata", "utils/residue_metadata.py"),
        asd=load("jamun.utils.average_squared_distance", "utils/average_squared_distance.py"),
        dist=load("jamun.distributions._distributions", "distributions/_distributions.py"),
        lr=load("jamun.lr_schedules._lr_schedules", "lr_schedules/_lr_schedules.py"),
    )
    return ref


# ----------------------------------------------------------------------------------------------------------------------
# recorded Gaussian draws (the reference calls torch.randn_like; the tests replay the same numbers into the kernels)
# ----------------------------------------------------------------------------------------------------------------------
class RecordedNoise:
    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.draws = []
        self._orig = torch.randn_like

    def __enter__(self):
        def randn_like(t, **k):
            r = torch.randn(t.shape, generator=self.gen, dtype=t.dtype)
            self.draws.append(r)
            return r

        torch.randn_like = randn_like
        return self

    def __exit__(self, *a):
        torch.randn_like = self._orig


def main():
    from toy_arch import TOY_W, toy_g

    ref = load_reference()
    out = {}
    gen = torch.Generator().manual_seed(2024)
    sizes = [22, 15, 9, 30, 2, 1, 57]
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    N = int(batch.shape[0])
    x = torch.randn(N, 3, generator=gen) * 0.4 + torch.randn(len(sizes), 3, generator=gen)[batch]
    out["sizes"] = np.array(sizes)

    # ---- 1. Kabsch (utils/align.py:9-56); the reference casts its one-hot to fp32, so it only runs in fp32
    y = x.clone()
    for c in range(len(sizes)):
        m = batch == c
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen))
        if torch.det(q) < 0:
            q[:, 0] = -q[:, 0]
        y[m] = x[m] @ q.T + torch.randn(3, generator=gen) + 0.05 * torch.randn(int(m.sum()), 3, generator=gen)
    out["kabsch_x"], out["kabsch_y"] = x.numpy(), y.numpy()
    out["kabsch_out_f32"] = ref.align.kabsch_algorithm(y, x, batch, len(sizes)).numpy()

    # ---- 2. BAOAB / ABOBA with a closed-form score (functional/_splitting.py:44-178)
    y0 = torch.randn(50, 3, generator=gen)
    kw = dict(delta=0.1, friction=0.7, M=2.0, inverse_temperature=1.3, score_fn_clip=2.5, steps=9, save_trajectory=True,
              save_every_n_steps=2, burn_in_steps=2)
    out["mcmc_y0"] = y0.numpy()
    out["mcmc_kwargs"] = np.array(repr(kw))
    for name in ("baoab", "aboba"):
        with RecordedNoise(7) as rec:
            yy, vv, yt, st = getattr(ref.splitting, name)(y0, lambda t: -3.0 * t, v_init="gaussian", **kw)
        out[f"{name}_noise"] = torch.stack(rec.draws).numpy()
        out[f"{name}_y"], out[f"{name}_v"], out[f"{name}_y_traj"], out[f"{name}_score_traj"] = (t.numpy() for t in (yy, vv, yt, st))

    # ---- 3. the real Denoiser around a closed-form network (model/denoiser.py)
    class ToyArch(torch.nn.Module):
        def forward(self, data, c_noise, effective_radial_cutoff):
            res = data.clone("pos")
            res.pos = toy_g(data.pos, c_noise, effective_radial_cutoff)
            return res

    asd, max_radius = 0.332, 1.0
    den = ref.denoiser.Denoiser(arch=ToyArch, optim=None, sigma_distribution=None, max_radius=max_radius,
                                average_squared_distance=asd, add_fixed_noise=False, add_fixed_ones=False,
                                align_noisy_input_during_training=True, align_noisy_input_during_evaluation=True,
                                mean_center=True, mirror_augmentation_rate=0.0, use_torch_compile=False)
    out["toy_w"] = TOY_W.numpy()
    out["asd"], out["max_radius"] = np.float64(asd), np.float64(max_radius)
    sigmas = [0.04, 0.1, 0.5]
    out["sigmas"] = np.array(sigmas)
    nf = []
    for s in sigmas:
        st = torch.as_tensor(s, dtype=torch.float32)
        c_in, c_skip, c_out, c_noise = den.normalization_factors(st, asd, 3)
        nf.append([float(c_in), float(c_skip), float(c_out), float(c_noise), float(den.effective_radial_cutoff(st)),
                   float(den.loss_weight(st, asd, 3))])
    out["normalization"] = np.array(nf, dtype=np.float64)  # fp32 values, widened

    def graph(pos):
        return Batch(pos=pos.clone(), batch=batch, edge_index=torch.zeros(2, 0, dtype=torch.long),
                     loss_weight=torch.linspace(0.5, 1.5, len(sizes)))

    yn = x + 0.04 * torch.randn(N, 3, generator=gen)
    out["den_y"] = yn.numpy()
    for k, s in enumerate(sigmas):
        with torch.no_grad():
            out[f"den_xhat_{k}"] = den.xhat(graph(yn), s).pos.numpy()
            out[f"den_score_{k}"] = den.score(graph(yn), s).numpy()
    out["mean_center_out"] = ref.mean_center.mean_center(graph(yn)).pos.numpy()

    # ---- 4. noise_and_denoise + compute_loss (denoiser.py:219-297): training-side tail incl. Kabsch alignment
    sig = 0.04
    with RecordedNoise(21) as rec:
        torch.manual_seed(0)  # the mirror-augmentation draw torch.rand(()) (rate 0: never mirrors)
        with torch.no_grad():
            xhat_g, y_g = den.noise_and_denoise(graph(x), sig, align_noisy_input=True)
            loss, aux = den.compute_loss(graph(x), xhat_g, torch.as_tensor(sig))
    out["train_x"], out["train_noise"] = x.numpy(), rec.draws[0].numpy()
    out["train_y_aligned"], out["train_xhat"] = y_g.pos.numpy(), xhat_g.pos.numpy()
    out["train_loss"] = loss.numpy()
    out["train_raw"], out["train_rmsd"] = aux["raw_coordinate_loss"].numpy(), aux["scaled_rmsd"].numpy()
    out["train_loss_weight"] = graph(x).loss_weight.numpy()

    # ---- 5. the walk-jump shell: SingleMeasurementSampler.sample over ModelSamplingWrapper over the Denoiser
    steps = 7
    wkw = dict(delta=0.04, friction=1.0, M=1.0, inverse_temperature=1.0, score_fn_clip=100.0, steps=steps, save_trajectory=True,
               save_every_n_steps=1, burn_in_steps=0)
    wrapped = ref.sampling_wrapper.ModelSamplingWrapper(den, graph(x), sig)
    mcmc = lambda y, score_fn, v_init: ref.splitting.baoab(y, score_fn, v_init=v_init, **wkw)  # noqa: E731
    sampler = ref.single.SingleMeasurementSampler(mcmc=mcmc, sigma=sig)
    with RecordedNoise(33) as rec:
        with torch.no_grad():
            res = sampler.sample(wrapped, y_init=yn.clone(), v_init="gaussian")
    out["walk_kwargs"] = np.array(repr(wkw))
    out["walk_noise"] = torch.stack(rec.draws).numpy()
    for key in ("xhat", "y", "v", "xhat_traj", "y_traj", "score_traj", "t_traj"):
        out[f"walk_{key}"] = res[key].numpy()
    assert res["sample"] is res["xhat"]
    # the same walk with clipping active and beta != 1 (score norms here are ~1e2)
    wkw2 = dict(wkw, score_fn_clip=30.0, inverse_temperature=0.8, M=1.5, friction=0.5, save_every_n_steps=2, burn_in_steps=1)
    mcmc2 = lambda y, score_fn, v_init: ref.splitting.baoab(y, score_fn, v_init=v_init, **wkw2)  # noqa: E731
    with RecordedNoise(34) as rec:
        with torch.no_grad():
            res = ref.single.SingleMeasurementSampler(mcmc=mcmc2, sigma=sig).sample(wrapped, y_init=yn.clone(), v_init="gaussian")
    out["walk2_kwargs"] = np.array(repr(wkw2))
    out["walk2_noise"] = torch.stack(rec.draws).numpy()
    for key in ("xhat", "y", "v", "xhat_traj", "y_traj", "score_traj"):
        out[f"walk2_{key}"] = res[key].numpy()

    # ---- 6. noise-conditioning MLPs and the atom embedding (pure torch inside the reference modules)
    torch.manual_seed(5)
    ncs = ref.noise.NoiseConditionalScaling(Irreps("120x0e + 32x1e"))
    skip = ref.noise.NoiseConditionalSkipConnection(Irreps("120x0e + 32x1e"))
    with torch.no_grad():
        for m in (ncs.scale_predictor, skip.weights.scale_predictor):
            m[-1].weight.normal_(0, 0.1)
            m[-1].bias.normal_(1, 0.1)
        c_noise = torch.tensor([math.log(0.04) / 4], dtype=torch.float32)
        out["ncs_scales"] = ncs.scale_predictor(c_noise).numpy()
        out["skip_weights"] = torch.sigmoid(skip.weights.scale_predictor(c_noise)).numpy()
    for k, v in ncs.state_dict().items():
        out[f"ncs_sd.{k}"] = v.numpy()
    for k, v in skip.state_dict().items():
        out[f"skip_sd.{k}"] = v.numpy()
    emb = ref.embed.AtomEmbeddingWithResidueInformation(8, 8, 32, 8, use_residue_sequence_index=False)
    idx = {k: torch.randint(0, hi, (N,), generator=gen) for k, hi in
           (("atom_type_index", 20), ("atom_code_index", 10), ("residue_code_index", 25), ("residue_sequence_index", 10))}
    with torch.no_grad():
        out["embed_out"] = emb(Data(**idx)).numpy()
    for k, v in emb.state_dict().items():
        out[f"embed_sd.{k}"] = v.numpy()
    for k, v in idx.items():
        out[f"embed_idx.{k}"] = v.numpy()

    # ---- 7. vocabulary (utils/residue_metadata.py): the order of these lists is the embedding-row order of released checkpoints
    M = ref.meta.ResidueMetadata
    out["meta_atom_types"], out["meta_atom_codes"], out["meta_residue_codes"] = (np.array(v) for v in (M.ATOM_TYPES, M.ATOM_CODES, M.RESIDUE_CODES))
    out["meta_aa3_keys"], out["meta_aa3_values"] = np.array(list(M.AA_3CODES.keys())), np.array(list(M.AA_3CODES.values()))
    probes = ["C", "O", "N", "F", "S", "H", "CA", "CB", "CG", "OXT", "ALA", "NME", "ACE", "HOH", "XYZ", ""]
    out["meta_probes"] = np.array(probes)
    out["meta_enc_type"] = np.array([ref.meta.encode_atom_type(p) for p in probes])
    out["meta_enc_code"] = np.array([ref.meta.encode_atom_code(p) for p in probes])
    out["meta_enc_res"] = np.array([ref.meta.encode_residue(p) for p in probes])
    peptides = ["AKT", "ala_LYS_thr", "W", "GLY", "ACDEFGHIKLMNPQRSTVWY"]
    out["meta_peptides"] = np.array(peptides)
    out["meta_three"] = np.array([ref.meta.convert_to_three_letter_codes(p) for p in peptides])
    out["meta_one"] = np.array([ref.meta.convert_to_one_letter_codes(p) for p in peptides])

    # ---- 8. average squared distance (utils/average_squared_distance.py:153-178), per chain, with and without a cut-off
    xs = x.numpy().astype(np.float64)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    cuts = [None, 0.4, 1.0]
    out["asd_cutoffs"] = np.array([-1.0 if c is None else c for c in cuts])
    out["asd_values"] = np.array([[ref.asd.compute_average_squared_distance(xs[offs[c]:offs[c + 1]], cutoff=cut) if sizes[c] > 1 else np.nan
                                   for c in range(len(sizes))] for cut in cuts])

    # ---- 9. sigma / measurement distributions (distributions/_distributions.py): samples under torch.manual_seed(123)
    D = ref.dist
    makers = {"constant": lambda m: m.ConstantSigma(0.04), "uniform": lambda m: m.UniformSigma(0.5, 0.01),
              "exponential": lambda m: m.ExponentialSigma(50.0, 1e-2), "lognormal": lambda m: m.ClippedLogNormalSigma(-1.2, 1.5, 2.0),
              "uniform_plus_normal": lambda m: m.UniformPlusNormal(0.3, (4, 3)),
              "uniform_measurement": lambda m: m.UniformMeasurement(0.5, 4),
              "weighted_measurement": lambda m: m.WeightedMeasurement(0.5, torch.tensor([0.1, 0.2, 0.3, 0.4]))}
    for name, mk in makers.items():
        torch.manual_seed(123)
        d = mk(D)
        out[f"dist_{name}"] = torch.stack([d.sample() for _ in range(5)] + [d.sample((3,))[i] for i in range(3)]).numpy()
    out["dist_measurement_mean"] = D.WeightedMeasurement(0.5, torch.tensor([0.1, 0.2, 0.3, 0.4])).mean.numpy()

    # ---- 10. LR multipliers (lr_schedules/_lr_schedules.py)
    steps_lr = [0, 1, 10, 99, 100, 101, 500, 1000, 5000]
    out["lr_steps"] = np.array(steps_lr)
    out["lr_warmup_decay"] = np.array([ref.lr.linear_warmup_linear_decay_lr_lambda(s_, num_warmup_steps=100, num_training_steps=1000) for s_ in steps_lr])
    out["lr_warmup_plateau"] = np.array([ref.lr.linear_warmup_plateau_lr_lambda(s_, num_warmup_steps=100, start_factor=0.1, end_factor=0.8) for s_ in steps_lr])
    out["lr_linear"] = np.array([ref.lr.linear(s_, start_factor=0.2, slope=-1e-3) for s_ in steps_lr])

    path = os.path.join(HERE, "reference_exec.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
