"""This library's kernels against vectors produced by executing the reference's own source files
(tests/golden/reference_exec.npz; generator tests/golden/make_reference_golden.py).  The network is the closed-form toy of
tests/toy_arch.py on both sides (evaluated with torch here as scaffolding), so these tests hold the kernels *around* the
network -- center_scale, walk_step (xhat / score / clip / BAOAB / jump), the generic BAOAB / ABOBA kernels, Kabsch, the
loss, the noise MLP and the embedding gather -- to the reference code itself rather than to the in-repo oracle."""
import ast
import math
import os

import numpy as np
import pytest
import torch

from toy_arch import TOY_W, toy_g

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_exec.npz")
SIGMA = 0.04


@pytest.fixture(scope="module")
def gold():
    z = np.load(GOLD)
    assert np.array_equal(z["toy_w"], TOY_W.numpy())
    return z


def C(a):
    return torch.from_numpy(np.asarray(a)).cuda().contiguous()


def _chains(z):
    sizes = [int(s) for s in z["sizes"]]
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    chain_ptr = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device="cuda")
    return sizes, batch.cuda(), chain_ptr


def _close(got, want, rtol, atol, what):
    got, want = got.detach().cpu(), torch.from_numpy(np.asarray(want))
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert torch.allclose(got, want, rtol=rtol, atol=atol), (what, (got - want).abs().max().item())


def test_kabsch_kernel_equals_reference_execution(gold):
    from jamun_b200.utils import kabsch_algorithm

    sizes, batch, _ = _chains(gold)
    got = kabsch_algorithm(C(gold["kabsch_y"]), C(gold["kabsch_x"]), batch, len(sizes))
    _close(got, gold["kabsch_out_f32"], 1e-4, 2e-5, "kabsch")


def test_generic_integrators_equal_reference_execution(gold):
    """BAOAB / ABOBA kernels behind the reference's mcmc(y, score_fn) protocol vs functional/_splitting.py itself."""
    from jamun_b200.sampling.mcmc import ABOBA, BAOAB

    kw = ast.literal_eval(str(gold["mcmc_kwargs"]))
    for cls, name in ((BAOAB, "baoab"), (ABOBA, "aboba")):
        got = cls(**kw)(C(gold["mcmc_y0"]), lambda y: -3.0 * y, v_init="gaussian", noise=C(gold[f"{name}_noise"]))
        for a, key in zip(got, ("y", "v", "y_traj", "score_traj")):
            _close(a, gold[f"{name}_{key}"], 1e-5, 1e-6, (name, key))


@pytest.mark.parametrize("tag", ["walk", "walk2"])
def test_fused_walk_kernels_equal_reference_walk_jump(gold, tag):
    """center_scale + walk_step driven exactly as fused_baoab drives them (first/last flags, trajectory slots, kept
    initial score), with the toy network in place of E3Conv, against SingleMeasurementSampler.sample of the reference."""
    from jamun_b200 import ops
    from jamun_b200.model.denoiser import SigmaContext
    from jamun_b200.sampling.mcmc.functional._splitting import _saves, _walk_params

    kw = ast.literal_eval(str(gold[f"{tag}_kwargs"]))
    steps, every, burn = kw["steps"], kw["save_every_n_steps"], kw["burn_in_steps"]
    _, _, chain_ptr = _chains(gold)
    ctx = SigmaContext(SIGMA, float(gold["asd"]), float(gold["max_radius"]))
    noise = C(gold[f"{tag}_noise"])
    y = C(gold["den_y"]).clone()
    N = y.shape[0]
    v = (math.sqrt(1.0 / kw["M"]) * noise[0]).contiguous()
    prm = _walk_params(kw["delta"], kw["friction"], kw["M"], kw["inverse_temperature"], kw["score_fn_clip"])
    prm.c_in, prm.c_skip, prm.c_out, prm.sigma2, prm.center = ctx.c_in, ctx.c_skip, ctx.c_out, ctx.sigma2, 1
    saved = [i for i in range(steps) if _saves(i, True, every, burn)]
    slot = {i: k for k, i in enumerate(saved)}
    off = int(0 not in slot)
    f32 = dict(dtype=torch.float32, device="cuda")
    y_traj, xhat_traj = torch.empty(len(saved), N, 3, **f32), torch.empty(len(saved), N, 3, **f32)
    score_traj = torch.empty(len(saved) + off, N, 3, **f32)
    ybar, p, xhat, score = (torch.empty(N, 3, **f32) for _ in range(4))
    ops.center_scale(y, chain_ptr, ctx.c_in, ybar, p, center=True)
    for i in range(steps):
        g = toy_g(p, ctx.c_noise, ctx.r_cut).contiguous()  # scaffolding: the network stand-in
        prm.first, prm.last, prm.step = int(i == 0), int(i == steps - 1), i + 1
        k = slot.get(i)
        ty, tx = (y_traj[k], xhat_traj[k]) if k is not None else (None, None)
        ts = score_traj[k + off] if k is not None else (score_traj[0] if i == 0 else None)
        nz = noise[i + 1].contiguous() if i + 1 < noise.shape[0] else None
        ops.walk_step(y, v, ybar, p, g, chain_ptr, prm, nz, xhat, score, ty, tx, ts)
    tol_s = 1e-5 / SIGMA ** 2
    _close(y, gold[f"{tag}_y"], 1e-4, 1e-5, "y")
    _close(v, gold[f"{tag}_v"], 1e-4, 1e-5, "v")
    _close(xhat, gold[f"{tag}_xhat"], 1e-4, 1e-5, "xhat")
    _close(y_traj, gold[f"{tag}_y_traj"], 1e-4, 1e-5, "y_traj")
    _close(xhat_traj, gold[f"{tag}_xhat_traj"], 1e-4, 1e-5, "xhat_traj")
    _close(score_traj, gold[f"{tag}_score_traj"], 1e-4, tol_s, "score_traj")
    want = torch.from_numpy(gold[f"{tag}_score_traj"])
    rel = ((score_traj.cpu() - want).norm() / want.norm()).item()
    assert rel <= 1e-4, rel


def test_xhat_normalized_call_path_equals_reference(gold):
    """Denoiser.xhat_normalized -> add_edges (radius kernel) -> g(y_scaled, c_noise, r_cut) with a foreign network: the
    arguments the network receives and the c_skip / c_out combination, for three noise levels."""
    import jamun_b200 as J
    from jamun_b200 import data, synthetic, utils

    class ToyArch(torch.nn.Module):
        def forward(self, graph, c_noise, r_cut):
            out = graph.clone("pos")
            out.pos = toy_g(graph.pos, c_noise, r_cut)
            return out

    sizes, _, _ = _chains(gold)
    den = J.default_denoiser(arch=ToyArch, max_radius=float(gold["max_radius"]), average_squared_distance=float(gold["asd"]))
    t = synthetic.make_tensors(sizes)
    batch = data.Batch.from_tensors(t).to("cuda")
    batch.pos = C(gold["den_y"])
    _close(utils.mean_center(batch.clone("pos")).pos, gold["mean_center_out"], 0, 1e-6, "mean_center")
    for k, s in enumerate(gold["sigmas"]):
        s = float(s)
        yb = utils.mean_center(batch.clone("pos"))
        xh = utils.mean_center(den.xhat_normalized(yb, s)).pos
        _close(xh, gold[f"den_xhat_{k}"], 1e-4, 1e-5, ("xhat", s))
        ctx = den.sigma_context(s)
        nf = gold["normalization"][k]
        got = [ctx.c_in, ctx.c_skip, ctx.c_out, ctx.c_noise]
        assert got == [float(w) for w in nf[:4]], (got, nf)
        assert abs(ctx.r_cut - float(nf[4]) / float(nf[0])) <= 1e-6 * ctx.r_cut


def test_training_tail_equals_reference(gold, monkeypatch):
    """mean_center -> add_noise -> mean_center -> Kabsch alignment (kernel), then the loss kernel on the reference's xhat."""
    import jamun_b200 as J
    from jamun_b200 import data, synthetic, utils

    sizes, _, _ = _chains(gold)
    den = J.default_denoiser().to("cuda")
    den.average_squared_distance, den.max_radius = float(gold["asd"]), float(gold["max_radius"])
    t = synthetic.make_tensors(sizes)
    t["loss_weight"] = torch.from_numpy(gold["train_loss_weight"])
    x = data.Batch.from_tensors(t).to("cuda")
    x.pos = C(gold["train_x"])
    nz = C(gold["train_noise"])
    monkeypatch.setattr(torch, "randn_like", lambda ten, **k: nz.to(ten.dtype))
    with torch.no_grad():
        xc = utils.mean_center(x)
        y = utils.mean_center(den.add_noise(xc, torch.as_tensor(SIGMA)))
        y = utils.align_A_to_B_batched(y, xc)
    monkeypatch.undo()
    _close(y.pos, gold["train_y_aligned"], 1e-4, 2e-5, "aligned y")
    xh = x.clone("pos")
    xh.pos = C(gold["train_xhat"])
    loss, aux = den.compute_loss(x, xh, SIGMA)
    _close(loss, gold["train_loss"], 1e-4, 1e-6, "loss")
    _close(aux["raw_coordinate_loss"], gold["train_raw"], 1e-4, 1e-8, "raw")
    _close(aux["scaled_rmsd"], gold["train_rmsd"], 1e-4, 1e-6, "rmsd")


def test_noise_mlp_and_embedding_kernels_equal_reference(gold):
    from jamun_b200 import data
    from jamun_b200.model.atom_embedding import AtomEmbeddingWithResidueInformation
    from jamun_b200.model.noise_conditioning import NoiseConditionalScaling, NoiseConditionalSkipConnection

    def sd(prefix):
        return {k[len(prefix):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith(prefix)}

    ncs = NoiseConditionalScaling("120x0e + 32x1e")
    ncs.load_state_dict(sd("ncs_sd."), strict=True)
    skip = NoiseConditionalSkipConnection("120x0e + 32x1e")
    skip.load_state_dict(sd("skip_sd."), strict=True)
    ncs, skip = ncs.cuda(), skip.cuda()
    c = math.log(SIGMA) / 4
    _close(ncs.scales(c), gold["ncs_scales"], 1e-5, 1e-6, "scales")
    _close(skip.weights.scales(c, sigmoid=True), gold["skip_weights"], 1e-5, 1e-6, "skip weights")
    emb = AtomEmbeddingWithResidueInformation(8, 8, 32, 8, use_residue_sequence_index=False)
    emb.load_state_dict(sd("embed_sd."), strict=True)
    emb = emb.cuda()
    idx = {k: v.cuda() for k, v in sd("embed_idx.").items()}
    got = emb(data.Data(**idx))
    _close(got, gold["embed_out"], 0, 0, "embedding")


def test_average_squared_distance_kernel_equals_reference(gold):
    """jamun_avg_sq_dist against utils/average_squared_distance.py (numpy, executed from the reference): per chain, with and
    without a cut-off, and the batched mean over graphs."""
    from jamun_b200.utils import compute_average_squared_distance

    sizes, _, chain_ptr = _chains(gold)
    x = C(gold["train_x"])
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for cut, want in zip(gold["asd_cutoffs"], gold["asd_values"]):
        cutoff = None if cut < 0 else float(cut)
        per_chain = []
        for c, n in enumerate(sizes):
            if n < 2:
                continue
            got = compute_average_squared_distance(x[offs[c]:offs[c + 1]], cutoff)
            assert abs(got - want[c]) <= 1e-5 * max(1.0, abs(want[c])), (cut, c, got, want[c])
            per_chain.append(want[c])
        # batched form: the mean over graphs (a single-atom graph contributes 0 pairs -> 0, as max(count, 1) in the kernel)
        got_all = compute_average_squared_distance(x, cutoff, chain_ptr=chain_ptr)
        assert abs(got_all - sum(per_chain) / len(sizes)) <= 1e-5
