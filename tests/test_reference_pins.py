"""The oracle against vectors produced by executing the reference's own source files
(tests/golden/reference_exec.npz, written by tests/golden/make_reference_golden.py in the build container).

These pin everything *around* the network -- normalisation factors, cut-off, centring, xhat / score tail, clip, BAOAB /
ABOBA, the walk-jump shell with its redundant jump pass, add_noise, Kabsch, the loss, the noise-conditioning MLPs and the
atom embedding -- to the reference code itself.  The e3nn arithmetic inside the network is not executable here and stays
pinned by substitutes only."""
import ast
import math
import os

import numpy as np
import pytest
import torch

from oracle import jamun_oracle as O
from toy_arch import TOY_W, toy_g

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_exec.npz")


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    assert np.array_equal(z["toy_w"], TOY_W.numpy()), "tests/toy_arch.py changed: regenerate the fixture"
    return z


def T(a):
    return torch.from_numpy(np.asarray(a))


class ToyArch(torch.nn.Module):
    def forward(self, data, c_noise, r_cut):
        return toy_g(data.pos, c_noise, r_cut)


def oracle_denoiser(z):
    return O.Denoiser(arch=ToyArch, max_radius=float(z["max_radius"]), average_squared_distance=float(z["asd"]), mean_center=True)


def oracle_batch(z, pos):
    sizes = [int(s) for s in z["sizes"]]
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    zeros = torch.zeros(len(batch), dtype=torch.long)
    return O.OracleBatch(pos=pos, batch=batch, num_graphs=len(sizes), edge_index=torch.zeros(2, 0, dtype=torch.long),
                         atom_type_index=zeros, atom_code_index=zeros, residue_code_index=zeros, residue_sequence_index=zeros,
                         loss_weight=T(z["train_loss_weight"]))


def replay(noise):
    it = iter(T(noise))
    return lambda y: next(it)


def test_kabsch_equals_reference_execution(gold):
    b = oracle_batch(gold, T(gold["kabsch_x"]))
    got = O.kabsch_algorithm(T(gold["kabsch_y"]), T(gold["kabsch_x"]), b.batch, b.num_graphs)
    big = torch.bincount(b.batch)[b.batch] >= 3  # 1-/2-atom chains: rank-deficient covariance, R not unique
    assert torch.allclose(got[big], T(gold["kabsch_out_f32"])[big], rtol=1e-5, atol=2e-6)
    assert torch.allclose(got[~big], T(gold["kabsch_out_f32"])[~big], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("name", ["baoab", "aboba"])
def test_integrators_equal_reference_execution(gold, name):
    kw = ast.literal_eval(str(gold["mcmc_kwargs"]))
    fn = getattr(O, name)
    y, v, yt, st = fn(T(gold["mcmc_y0"]), lambda t: -3.0 * t, v_init="gaussian", noise_fn=replay(gold[f"{name}_noise"]), **kw)
    for got, key in ((y, "y"), (v, "v"), (yt, "y_traj"), (st, "score_traj")):
        want = T(gold[f"{name}_{key}"])
        assert got.shape == want.shape, key
        assert torch.equal(got, want), (name, key, (got - want).abs().max())  # same fp32 operations in the same order


def test_normalization_cutoff_and_loss_weight_equal_reference(gold):
    den = oracle_denoiser(gold)
    for s, want in zip(gold["sigmas"], gold["normalization"]):
        st = torch.as_tensor(float(s), dtype=torch.float32)
        c_in, c_skip, c_out, c_noise = den.normalization_factors(st, den.average_squared_distance, 3)
        got = [float(c_in), float(c_skip), float(c_out), float(c_noise), float(den.effective_radial_cutoff(st)), float(1 / c_out ** 2)]
        assert got == [float(w) for w in want], (s, got, want)


def test_xhat_score_and_centering_equal_reference(gold):
    den = oracle_denoiser(gold)
    y = T(gold["den_y"])
    b = oracle_batch(gold, y)
    assert torch.allclose(O.mean_center_pos(y, b.batch, b.num_graphs), T(gold["mean_center_out"]), rtol=0, atol=1e-7)
    for k, s in enumerate(gold["sigmas"]):
        xh = den.xhat(b, float(s))
        sc = den.score(b, float(s))
        assert torch.allclose(xh, T(gold[f"den_xhat_{k}"]), rtol=1e-6, atol=1e-6), (s, (xh - T(gold[f"den_xhat_{k}"])).abs().max())
        assert torch.allclose(sc, T(gold[f"den_score_{k}"]), rtol=1e-5, atol=1e-6 / float(s) ** 2)


def test_noise_align_and_loss_equal_reference(gold):
    den = oracle_denoiser(gold)
    b = oracle_batch(gold, T(gold["train_x"]))
    xhat, ypos = O.noise_and_denoise(den, b, 0.04, T(gold["train_noise"]), align_noisy_input=True)
    assert torch.allclose(ypos, T(gold["train_y_aligned"]), rtol=1e-5, atol=2e-6)
    assert torch.allclose(xhat, T(gold["train_xhat"]), rtol=1e-5, atol=2e-6)
    loss, aux = O.compute_loss(den, b, T(gold["train_xhat"]), 0.04)
    assert torch.allclose(loss, T(gold["train_loss"]), rtol=1e-5, atol=1e-7)
    assert torch.allclose(aux["raw_coordinate_loss"], T(gold["train_raw"]), rtol=1e-5, atol=1e-9)
    assert torch.allclose(aux["scaled_rmsd"], T(gold["train_rmsd"]), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("tag", ["walk", "walk2"])
def test_walk_jump_shell_equals_reference(gold, tag):
    """SingleMeasurementSampler.sample over ModelSamplingWrapper over Denoiser over baoab, executed from the reference's
    files, against O.walk_jump -- including the jump pass over every saved frame and the kept score of y_init."""
    den = oracle_denoiser(gold)
    kw = ast.literal_eval(str(gold[f"{tag}_kwargs"]))
    b = oracle_batch(gold, T(gold["train_x"]))
    out = O.walk_jump(den, b, T(gold["den_y"]), 0.04, mcmc=O.baoab, v_init="gaussian", noise_fn=replay(gold[f"{tag}_noise"]), **kw)
    for key in ("xhat", "y", "v", "xhat_traj", "y_traj", "score_traj"):
        want = T(gold[f"{tag}_{key}"])
        assert out[key].shape == want.shape, key
        tol = dict(rtol=1e-5, atol=1e-6 if "score" not in key else 1e-6 / 0.04 ** 2)
        assert torch.allclose(out[key], want, **tol), (key, (out[key] - want).abs().max())
    if tag == "walk":
        assert torch.equal(out["t_traj"], T(gold["walk_t_traj"]))
    # BAOAB: the jump is a by-product of the score -- xhat_traj == y_traj + sigma^2 * score_traj on the reference's own output
    yt, st, xt = (T(gold[f"{tag}_{k}"]) for k in ("y_traj", "score_traj", "xhat_traj"))
    off = st.shape[0] - yt.shape[0]  # score(y_init) is kept even when burn-in drops y_init
    assert torch.allclose(yt + 0.04 ** 2 * st[off:], xt, rtol=0, atol=5e-7)


def test_noise_conditioning_and_embedding_equal_reference(gold):
    ncs = O.NoiseConditionalScaling("120x0e + 32x1e")
    ncs.load_state_dict({k[len("ncs_sd."):]: T(gold[k]) for k in gold.files if k.startswith("ncs_sd.")}, strict=True)
    skip = O.NoiseConditionalSkipConnection("120x0e + 32x1e")
    skip.load_state_dict({k[len("skip_sd."):]: T(gold[k]) for k in gold.files if k.startswith("skip_sd.")}, strict=True)
    c_noise = torch.tensor([math.log(0.04) / 4], dtype=torch.float32)
    with torch.no_grad():
        assert torch.allclose(ncs.scale_predictor(c_noise), T(gold["ncs_scales"]), rtol=1e-6, atol=1e-7)
        assert torch.allclose(torch.sigmoid(skip.weights.scale_predictor(c_noise)), T(gold["skip_weights"]), rtol=1e-6, atol=1e-7)
    emb = O.AtomEmbeddingWithResidueInformation(8, 8, 32, 8, use_residue_sequence_index=False)
    emb.load_state_dict({k[len("embed_sd."):]: T(gold[k]) for k in gold.files if k.startswith("embed_sd.")}, strict=True)
    idx = {k[len("embed_idx."):]: T(gold[k]) for k in gold.files if k.startswith("embed_idx.")}
    n = len(idx["atom_type_index"])
    b = O.OracleBatch(pos=torch.zeros(n, 3), batch=torch.zeros(n, dtype=torch.long), num_graphs=1,
                      edge_index=torch.zeros(2, 0, dtype=torch.long), **idx)
    with torch.no_grad():
        assert torch.equal(emb(b), T(gold["embed_out"]))


def test_vocabulary_and_name_helpers_equal_reference(gold):
    """jamun_b200.utils.residue_metadata (host logic of the product, not the oracle) against the reference's own tables and
    helpers: the list order is the embedding-row order a released checkpoint learned."""
    from jamun_b200.utils import residue_metadata as RM

    M = RM.ResidueMetadata
    assert M.ATOM_TYPES == list(gold["meta_atom_types"]) and M.ATOM_CODES == list(gold["meta_atom_codes"])
    assert M.RESIDUE_CODES == list(gold["meta_residue_codes"])
    assert list(M.AA_3CODES.keys()) == list(gold["meta_aa3_keys"]) and list(M.AA_3CODES.values()) == list(gold["meta_aa3_values"])
    probes = [str(p) for p in gold["meta_probes"]]
    assert [RM.encode_atom_type(p) for p in probes] == list(gold["meta_enc_type"])
    assert [RM.encode_atom_code(p) for p in probes] == list(gold["meta_enc_code"])
    assert [RM.encode_residue(p) for p in probes] == list(gold["meta_enc_res"])
    peptides = [str(p) for p in gold["meta_peptides"]]
    assert [RM.convert_to_three_letter_codes(p) for p in peptides] == list(gold["meta_three"])
    assert [RM.convert_to_one_letter_codes(p) for p in peptides] == list(gold["meta_one"])
    for bad in ("B", "XYZ", "ALAA"):
        with pytest.raises(ValueError):
            RM.convert_to_three_letter_code(bad)


def test_sigma_distributions_and_lr_schedules_equal_reference(gold):
    """jamun_b200.distributions / jamun_b200.lr_schedules (host logic of the product) against the reference's own classes run
    under the same seed: every Hydra `sigma_distribution` target and the LambdaLR multipliers of `lr_scheduler_config`."""
    from jamun_b200 import distributions as D
    from jamun_b200 import lr_schedules as L

    makers = {"constant": lambda: D.ConstantSigma(0.04), "uniform": lambda: D.UniformSigma(0.5, 0.01),
              "exponential": lambda: D.ExponentialSigma(50.0, 1e-2), "lognormal": lambda: D.ClippedLogNormalSigma(-1.2, 1.5, 2.0),
              "uniform_plus_normal": lambda: D.UniformPlusNormal(0.3, (4, 3)),
              "uniform_measurement": lambda: D.UniformMeasurement(0.5, 4),
              "weighted_measurement": lambda: D.WeightedMeasurement(0.5, torch.tensor([0.1, 0.2, 0.3, 0.4]))}
    for name, mk in makers.items():
        torch.manual_seed(123)
        d = mk()
        got = torch.stack([d.sample() for _ in range(5)] + [d.sample((3,))[i] for i in range(3)])
        assert torch.equal(got, T(gold[f"dist_{name}"])), name
    mean = D.WeightedMeasurement(0.5, torch.tensor([0.1, 0.2, 0.3, 0.4])).mean
    assert torch.allclose(mean, T(gold["dist_measurement_mean"]), rtol=1e-6, atol=0)
    steps = [int(s) for s in gold["lr_steps"]]
    assert [L.linear_warmup_linear_decay_lr_lambda(s, num_warmup_steps=100, num_training_steps=1000) for s in steps] == list(gold["lr_warmup_decay"])
    assert [L.linear_warmup_plateau_lr_lambda(s, num_warmup_steps=100, start_factor=0.1, end_factor=0.8) for s in steps] == list(gold["lr_warmup_plateau"])
    assert [L.linear(s, start_factor=0.2, slope=-1e-3) for s in steps] == list(gold["lr_linear"])
