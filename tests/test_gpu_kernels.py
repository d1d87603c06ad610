"""GPU parity tests, kernel by kernel, through the C ABI (jamun_b200.ops -> libjamun_b200.so) against the oracle."""
import math

import pytest
import torch

import kernel_model as KM
from conftest import make_oracle_batch

pytestmark = pytest.mark.gpu

SIGMA = 0.04


def _setup(models, sizes, seed=0, dtype=torch.float32):
    import jamun_b200 as J
    from jamun_b200 import synthetic
    from oracle import jamun_oracle as O

    o32, o64, prod = models
    t = synthetic.make_tensors(sizes)
    gen = torch.Generator().manual_seed(seed)
    y = t["pos"] + SIGMA * torch.randn(t["pos"].shape, generator=gen)
    return o32, o64, prod, t, y


def test_library_loaded():
    from jamun_b200 import _lib

    assert _lib.lib().jamun_abi_version() == 1


@pytest.mark.parametrize("sizes", [[22, 15, 9, 30], [1], [2, 1, 57], [40] * 5, [200]])
def test_center_scale(sizes):
    from jamun_b200 import engine, ops, synthetic, data
    from oracle import jamun_oracle as O

    t = synthetic.make_tensors(sizes)
    y = (t["pos"] + 0.3).cuda()
    topo = engine.Topology(data.Batch.from_tensors(t), "cuda")
    ybar, p = ops.center_scale(y, topo.chain_ptr, 1.7)
    ref = O.mean_center_pos(y.cpu(), t["batch"], len(sizes))
    assert torch.allclose(ybar.cpu(), ref, atol=1e-6, rtol=0)
    assert torch.allclose(p.cpu(), ref * 1.7, atol=1e-6, rtol=0)


@pytest.mark.parametrize("sizes,cap", [([22, 15, 9, 30], 32), ([60, 57, 3, 1, 2], 32), ([150, 40], 32), ([150, 40], 8),
                                       ([150, 40], None), ([300], 32), ([1000] * 4, 32), ([3000], 32), ([300, 5, 70, 1, 2], 8)])
def test_radius_csr_bit_exact(sizes, cap):
    """Edge sets identical to the oracle's radius_graph + bonded concatenation (compared after sorting), and the
    within-row order is the documented one (radial ascending, then bonded)."""
    from jamun_b200 import engine, synthetic, data
    from oracle import jamun_oracle as O

    t = synthetic.make_tensors(sizes)
    pos = t["pos"]
    r = 0.5872642993927002
    topo = engine.Topology(data.Batch.from_tensors(t), "cuda", max_num_neighbors=cap)
    topo.build_csr(pos.cuda(), r)
    ei, bm = topo.edge_index()
    ref_rad = O.radius_graph(pos, r, t["batch"], cap)
    ref = torch.cat([ref_rad, t["edge_index"]], dim=1)
    ref_bm = torch.cat([torch.zeros(ref_rad.shape[1], dtype=torch.long), torch.ones(t["edge_index"].shape[1], dtype=torch.long)])
    assert ei.shape[1] == ref.shape[1]
    N = pos.shape[0]
    key = lambda e, m: torch.sort(e[1] * (2 * N) * 2 + e[0] * 2 + m).values  # noqa: E731
    assert torch.equal(key(ei.cpu(), bm.cpu()), key(ref, ref_bm))
    # receiver-sorted
    assert bool((ei[1][1:] >= ei[1][:-1]).all())
    # stable reorder of the oracle's list by receiver reproduces ours exactly (row order contract)
    order = torch.sort(ref[1], stable=True).indices
    assert torch.equal(ei.cpu(), ref[:, order])
    rowptr = topo.rowptr.cpu().long()
    assert torch.equal(rowptr[1:] - rowptr[:-1], torch.bincount(ref[1], minlength=N))


@pytest.mark.parametrize("sizes,cap", [([22, 15, 9, 30, 1, 2], 32), ([1000, 200, 64, 65], 32), ([500] * 3, 5)])
def test_radius_cell_list_equals_brute_force(sizes, cap, monkeypatch):
    """The cell-list search (long chains) and the ascending brute-force scan (short chains) emit the same CSR, bit for bit."""
    from jamun_b200 import data, engine, synthetic

    t = synthetic.make_tensors(sizes)
    out = {}
    for impl in ("brute", "cells"):
        monkeypatch.setenv("JAMUN_B200_RADIUS", impl)
        topo = engine.Topology(data.Batch.from_tensors(t), "cuda", max_num_neighbors=cap)
        topo.build_csr(t["pos"].cuda(), 0.5872642993927002)
        E = int(topo.rowptr[-1])
        out[impl] = (topo.rowptr.cpu(), topo.col[:E].cpu(), topo.edst[:E].cpu(), topo.ebond[:E].cpu())
    for a, b in zip(out["brute"], out["cells"]):
        assert torch.equal(a, b)


def test_radius_empty_and_single():
    from jamun_b200 import engine, synthetic, data

    t = synthetic.make_tensors([1, 1])
    topo = engine.Topology(data.Batch.from_tensors(t), "cuda")
    topo.build_csr(t["pos"].cuda(), 0.5)
    assert int(topo.rowptr[-1]) == 0


def _ref_layers(o64, t, y):
    """fp64 oracle intermediates for positions y."""
    from oracle import jamun_oracle as O

    b = make_oracle_batch(t, torch.float64)
    sig = torch.tensor(SIGMA, dtype=torch.float64)
    ybar = O.mean_center_pos(y.double(), b.batch, b.num_graphs)
    c_in, c_skip, c_out, c_noise = o64.normalization_factors(sig, 0.332)
    r_cut = o64.effective_radial_cutoff(sig) / c_in
    yg = o64.add_edges(b.with_pos(ybar), r_cut)
    ys = yg.with_pos(ybar * c_in)
    g, hidden = o64.g(ys, c_noise.unsqueeze(0), r_cut, return_hidden=True)
    return dict(ybar=ybar, yg=yg, ys=ys, g=g, hidden=hidden, r_cut=float(r_cut), c_in=float(c_in), c_noise=float(c_noise))


@pytest.mark.parametrize("sizes", [[22, 15, 9, 30], [57, 3, 1, 40, 40, 17, 64, 65]])
def test_network_layers_vs_oracle(models, sizes):
    """Every stage of E3Conv against the fp64 oracle: edge features, radial hidden, each block's output, head."""
    from jamun_b200 import engine, ops, data

    o32, o64, prod, t, y = _setup(models, sizes)
    ref = _ref_layers(o64, t, y)
    g = prod.arch_module
    topo = engine.Topology(data.Batch.from_tensors(t), "cuda")
    ctx = prod.sigma_context(SIGMA)
    assert abs(ctx.r_cut - 0.5872642993927002) < 1e-6
    plan = g.plan(ctx.c_noise, "cuda")
    ybar, p = ops.center_scale(y.cuda(), topo.chain_ptr, ctx.c_in)
    topo.build_csr(ybar, ctx.r_cut)
    ei, bm = topo.edge_index()
    # same edge multiset as the oracle at these positions
    src, dst = ref["yg"].edge_index
    order = torch.sort(dst, stable=True).indices
    assert torch.equal(ei.cpu(), torch.stack([src[order], dst[order]]))
    E = ei.shape[1]
    mu, step = plan.radial_grid(ctx.r_cut)
    ops.edge_geom(p, topo.rowptr, topo.col, topo.edst, mu, step, topo.rhat, topo.rb)
    rhat_ref, rb_ref = KM.edge_geom(ref["ys"].pos, src[order], dst[order], ref["r_cut"])
    assert torch.allclose(topo.rhat[:E, :3].cpu().double(), rhat_ref, atol=2e-6)
    assert torch.allclose(topo.rb[:E].cpu().double(), rb_ref, atol=2e-6)
    g_out = torch.empty_like(p)
    # run block by block, checking node features after each block
    engine.e3conv_forward(plan, topo, p, ctx.r_cut, g_out)
    N = p.shape[0]
    # re-run with per-block inspection
    x_in, x_res = topo.x0, None
    nb = len(plan.blocks)
    for l, b in enumerate(plan.blocks):
        ops.edge_radial_hidden(topo.rb, topo.ebond, topo.rowptr, b["w0r"], b["b0eff"], topo.h)
        ops.conv_fwd(x_in, b["s_in"], b["v_in"], topo.rowptr, topo.col, topo.h, topo.rhat, b["m0"], b["m1"], b["alpha0"],
                     b["alpha1"], topo.conv)
        x_new, x_scaled = topo.xa[l & 1], topo.xs[l & 1]
        ops.block_tail(topo.conv, x_in, b["s_in"], b["v_in"], x_res, b["wself_s"], b["wself_v"], b["wskip_s"], b["wskip_v"],
                       plan.skips[l - 1] if l > 0 else None, plan.scales[l] if l < nb - 1 else None, b["c_act"], b["c_gate"],
                       x_new, x_scaled if l < nb - 1 else None)
        got = ops.layout_from_soa(x_new, 120, 32).cpu().double()
        want = ref["hidden"][l]
        err = (got - want).abs().max().item()
        scale = want.abs().max().item()
        assert err <= 2e-5 * max(1.0, scale), f"block {l}: err {err} scale {scale}"
        x_in, x_res = x_scaled, x_new
    err = (g_out.cpu().double() - ref["g"]).abs().max().item()
    assert err <= 1e-5, f"g err {err}"


@pytest.mark.parametrize("sizes", [[22] * 8, [9, 29, 17, 12, 25], [57, 41, 36, 17], [300, 5]])
def test_xhat_score_parity(models, sizes):
    """north_star tolerance: denoised coordinates and scores within rtol 1e-4 / atol 1e-5 (fp32), teacher-forced."""
    from jamun_b200 import data
    from oracle import jamun_oracle as O

    o32, o64, prod, t, y = _setup(models, sizes, seed=3)
    batch = data.Batch.from_tensors(t).to("cuda")
    yb = batch.clone("pos")
    yb.pos = y.cuda()
    xhat = prod.xhat(yb, SIGMA).pos.cpu()
    score = prod.score(yb, SIGMA).cpu()
    ob = make_oracle_batch(t)
    with torch.no_grad():
        xref = o32.xhat(ob.with_pos(y), SIGMA)
        sref = o32.score(ob.with_pos(y), SIGMA)
        xref64 = o64.xhat(make_oracle_batch(t, torch.float64).with_pos(y.double()), SIGMA)
    assert torch.allclose(xhat, xref, rtol=1e-4, atol=1e-5), (xhat - xref).abs().max()
    assert torch.allclose(xhat.double(), xref64, rtol=1e-4, atol=1e-5)
    assert torch.allclose(score, sref, rtol=1e-4, atol=1e-5 / SIGMA ** 2), (score - sref).abs().max()


def test_output_gain_zero_identity(models):
    """A freshly initialised model returns mean_center(c_skip * mean_center(y))."""
    import jamun_b200 as J
    from jamun_b200 import data, synthetic
    from oracle import jamun_oracle as O

    torch.manual_seed(1)
    m = J.default_denoiser().cuda()
    t = synthetic.make_tensors([22, 9])
    b = data.Batch.from_tensors(t).to("cuda")
    x = m.xhat(b, SIGMA).pos.cpu()
    yb = O.mean_center_pos(t["pos"], t["batch"], 2)
    c_skip = float(m.sigma_context(SIGMA).c_skip)
    assert torch.allclose(x, O.mean_center_pos(c_skip * yb, t["batch"], 2), atol=1e-6)


def test_equivariance_and_chain_independence(models):
    from jamun_b200 import data, synthetic

    o32, o64, prod, t, y = _setup(models, [22, 15, 9, 30], seed=5)
    batch = data.Batch.from_tensors(t).to("cuda")

    def run(pos):
        yb = batch.clone("pos")
        yb.pos = pos.cuda().contiguous()
        return prod.xhat(yb, SIGMA).pos.cpu()

    x = run(y)
    Q, _ = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(0)))
    if torch.det(Q) < 0:
        Q[:, 0] *= -1
    xr = run(y @ Q.T + torch.tensor([0.3, -0.2, 0.1]))
    assert torch.allclose(xr, x @ Q.T, atol=2e-5)
    # perturbing chain 0 leaves the other chains' outputs bit-identical
    y2 = y.clone()
    y2[:22] += 0.01
    x2 = run(y2)
    assert torch.equal(x2[22:], x[22:])


def test_walk_step_matches_oracle_baoab(models):
    """Fused walk (K1..K6 per step) against the oracle's Python BAOAB with the same supplied noise, saved every step."""
    from jamun_b200 import data, utils
    from jamun_b200.sampling.mcmc import BAOAB
    from jamun_b200.sampling.walkjump import SingleMeasurementSampler
    from oracle import jamun_oracle as O

    o32, o64, prod, t, y = _setup(models, [22, 15, 9], seed=7)
    steps = 6
    gen = torch.Generator().manual_seed(11)
    noise = torch.randn(steps, y.shape[0], 3, generator=gen)
    kw = dict(delta=0.04, friction=1.0, M=1.0, inverse_temperature=1.0, score_fn_clip=100.0)
    it = iter(noise)
    ob = make_oracle_batch(t)
    with torch.no_grad():
        ref = O.walk_jump(o32, ob, y, SIGMA, mcmc=O.baoab, v_init="gaussian", steps=steps, save_trajectory=True,
                          noise_fn=lambda yy: next(it), **kw)
    batch = data.Batch.from_tensors(t).to("cuda")
    wrapped = utils.ModelSamplingWrapper(prod, batch, SIGMA)
    sampler = SingleMeasurementSampler(BAOAB(steps=steps, save_trajectory=True, **kw), SIGMA)
    out = sampler.sample(wrapped, y_init=y.cuda(), v_init="gaussian", noise=noise.cuda())
    for key in ("y", "v", "xhat", "y_traj", "xhat_traj", "score_traj"):
        got, want = out[key].cpu(), ref[key]
        assert got.shape == want.shape, key
        # free-running: both walks integrate their own scores for 6 steps, so per-step differences compound (x2 allowance)
        tol = dict(rtol=2e-4, atol=2e-5 if "score" not in key else 2e-5 / SIGMA ** 2)
        assert torch.allclose(got, want, **tol), (key, (got - want).abs().max())
    assert out["sample"] is out["xhat"]
    # teacher-forced, per step, at north_star's tolerance: the kernels evaluated at the oracle's own y_t
    for i in range(steps):
        yb = batch.clone("pos")
        yb.pos = ref["y_traj"][i].cuda().contiguous()
        xh = prod.xhat(yb, SIGMA).pos.cpu()
        sc = prod.score(yb, SIGMA).cpu()
        assert torch.allclose(xh, ref["xhat_traj"][i], rtol=1e-4, atol=1e-5), (i, (xh - ref["xhat_traj"][i]).abs().max())
        sref = ref["score_traj"][i]
        # score = (xhat - y) / sigma^2: an absolute error of 1e-5 on xhat is 1e-5 / sigma^2 on the score
        assert torch.allclose(sc, sref, rtol=1e-4, atol=1e-5 / SIGMA ** 2), (i, (sc - sref).abs().max())
        rel = ((sc - sref).norm() / sref.norm()).item()
        assert rel <= 1e-4, f"step {i}: relative score error {rel:.2e}"


def test_generic_baoab_and_aboba_protocol():
    """The reference's mcmc(y, score_fn) protocol with a closed-form score (harmonic well): kernels vs oracle loops."""
    from jamun_b200.sampling.mcmc import ABOBA, BAOAB
    from oracle import jamun_oracle as O

    gen = torch.Generator().manual_seed(0)
    y0 = torch.randn(50, 3, generator=gen)
    steps = 9
    noise = torch.randn(steps, 50, 3, generator=gen)
    score = lambda y: -3.0 * y  # noqa: E731
    kw = dict(delta=0.1, friction=0.7, M=2.0, inverse_temperature=1.3, score_fn_clip=2.5, steps=steps, save_trajectory=True,
              save_every_n_steps=2, burn_in_steps=2)
    for cls, fn in ((BAOAB, O.baoab), (ABOBA, O.aboba)):
        it = iter(noise)
        ref = fn(y0, score, v_init="gaussian", noise_fn=lambda yy: next(it), **kw)
        got = cls(**kw)(y0.cuda(), score, v_init="gaussian", noise=noise.cuda())
        for a, b in zip(got, ref):
            assert a.shape == b.shape
            assert torch.allclose(a.cpu(), b, rtol=1e-5, atol=1e-6), cls
    with pytest.raises(RuntimeError):
        BAOAB(v_init="nope")


def test_philox_normals_statistics():
    from jamun_b200 import ops

    out = torch.empty(200_000, 3, device="cuda")
    ops.gaussian_axpy(None, 0.0, 1.0, None, 1234, 1, out)
    out2 = torch.empty_like(out)
    ops.gaussian_axpy(None, 0.0, 1.0, None, 1234, 2, out2)
    z = out.flatten()
    assert abs(z.mean().item()) < 0.01 and abs(z.std().item() - 1) < 0.01
    assert abs((z ** 4).mean().item() - 3.0) < 0.1
    assert abs(torch.corrcoef(torch.stack([out.flatten(), out2.flatten()]))[0, 1].item()) < 0.01
    assert not torch.equal(out, out2)


def test_no_cpu_fallback():
    from jamun_b200 import ops

    with pytest.raises(RuntimeError):
        ops.center_scale(torch.zeros(3, 3), torch.tensor([0, 3], dtype=torch.int32), 1.0)


def test_cuda_graph_replay_is_bit_identical_to_eager(models):
    """Steady-state steps replayed from a CUDA graph (device-side step counter / trajectory slot) == eager launches."""
    import itertools

    from jamun_b200 import data, utils
    from jamun_b200.sampling.mcmc.functional import _splitting, fused_baoab

    o32, o64, prod, t, y = _setup(models, [22, 15, 9, 30, 12], seed=9)
    batch = data.Batch.from_tensors(t).to("cuda")
    topo = prod.topology_for(batch)
    kw = dict(steps=9, v_init="gaussian", save_trajectory=True, save_every_n_steps=2, burn_in_steps=3, delta=0.04, friction=1.0,
              M=1.0, inverse_temperature=1.0, score_fn_clip=100.0)
    outs = []
    for use_graph in (False, True, True):
        torch.manual_seed(123)
        _splitting._call_counter = itertools.count(1)
        outs.append(fused_baoab(prod, topo, y.cuda(), SIGMA, use_cuda_graph=use_graph, **kw))
    for key in ("y", "v", "xhat", "y_traj", "xhat_traj", "score_traj"):
        assert outs[0][key].shape == outs[1][key].shape
        assert torch.equal(outs[0][key], outs[1][key]), key
        assert torch.equal(outs[0][key], outs[2][key]), key  # cached graph, second use
    assert outs[0]["y_traj"].shape[0] == 3 and outs[0]["score_traj"].shape[0] == 4  # frames 4,6,8 (+ the initial score)


def test_module_level_forwards_match_oracle_modules(models):
    """The reference's plug-in seams at module granularity: Conv.forward / apply_per_edge, ConvBlock, o3.Linear stand-in,
    Gate, EquivariantMLP, NoiseConditionalScaling/SkipConnection, atom embedder -- same signatures, e3nn layouts."""
    from jamun_b200 import synthetic

    o32, o64, prod = models
    g, og = prod.arch_module, o32.g
    t = synthetic.make_tensors([14, 9, 20])
    N = t["pos"].shape[0]
    gen = torch.Generator().manual_seed(4)
    ob = make_oracle_batch(t)
    ei = O_radius(t)
    E = ei.shape[1]
    bm = torch.zeros(E, dtype=torch.long)
    ea, sh = og.edge_features(t["pos"] * 1.7, ei, bm, 0.587)
    x = torch.randn(N, 216, generator=gen)
    with torch.no_grad():
        layer, olayer = g.layers[1], og.layers[1]
        want = olayer.gated_conv.f.f(x, ei, ea, sh)
        got = layer.conv(x.cuda(), ei.cuda(), ea.cuda(), sh.cuda()).cpu()
        assert torch.allclose(got, want, rtol=1e-4, atol=1e-5), (got - want).abs().max()
        want_e = olayer.gated_conv.f.f.tp(x[ei[0]], sh, olayer.gated_conv.f.f.radial_nn(ea))
        got_e = layer.conv.apply_per_edge(x[ei[0]].cuda(), ea.cuda(), sh.cuda()).cpu()
        assert torch.allclose(got_e, want_e, rtol=1e-4, atol=1e-5)
        assert torch.allclose(layer(x.cuda(), ei.cuda(), ea.cuda(), sh.cuda()).cpu(), olayer(x, ei, ea, sh), rtol=1e-4, atol=2e-5)
        assert torch.allclose(g.output_head(x.cuda()).cpu(), og.output_head(x), rtol=1e-4, atol=1e-5)
        lin, olin = layer.gated_conv.self_interaction, olayer.gated_conv.self_interaction
        assert torch.allclose(lin(x.cuda()).cpu(), olin(x), rtol=1e-5, atol=1e-5)
        c = torch.tensor([-0.8047])
        assert torch.allclose(g.noise_scalings[0](x.cuda(), c).cpu(), og.noise_scalings[0](x, c), rtol=1e-5, atol=1e-6)
        assert torch.allclose(g.skip_connections[0](x.cuda(), 2 * x.cuda(), c).cpu(), og.skip_connections[0](x, 2 * x, c), rtol=1e-5, atol=1e-6)
        from jamun_b200 import data
        b = data.Batch.from_tensors(t).to("cuda")
        assert torch.allclose(g.atom_embedder(b).cpu(), og.atom_embedder(ob), atol=0)
        x56 = torch.randn(N, 56, generator=gen)
        assert torch.allclose(g.initial_projector(x56.cuda(), ei.cuda(), ea.cuda(), sh.cuda()).cpu(),
                              og.initial_projector(x56, ei, ea, sh), rtol=1e-4, atol=2e-5)


def O_radius(t):
    from oracle import jamun_oracle as O

    return torch.cat([O.radius_graph(t["pos"], 0.587, t["batch"], 32), t["edge_index"]], dim=1)


def test_reference_call_path_xhat_normalized_and_explicit_edges(models):
    """Denoiser.xhat_normalized -> add_edges -> g(y_scaled, c_noise, r_cut) (the reference's own call sequence), and
    E3Conv.forward fed an explicit edge_index/bond_mask, both agree with the fused denoise_positions path."""
    from jamun_b200 import data, utils

    o32, o64, prod, t, y = _setup(models, [22, 15, 9], seed=2)
    batch = data.Batch.from_tensors(t).to("cuda")
    yb = utils.mean_center(batch.clone("pos"))
    yb2 = batch.clone("pos")
    yb2.pos = y.cuda()
    yb2 = utils.mean_center(yb2)
    fused = prod.xhat(yb2, SIGMA).pos
    xn = prod.xhat_normalized(yb2, SIGMA)
    assert torch.allclose(utils.mean_center(xn).pos, fused, atol=1e-6)
    # explicit edge list (materialised from the CSR), fresh graph object without a cached topology/CSR
    ctx = prod.sigma_context(SIGMA)
    yb3 = prod.add_edges(yb2.clone("pos"), ctx.r_cut, materialize=True)
    fresh = data.Batch.from_tensors(t).to("cuda")
    fresh.pos = yb2.pos * ctx.c_in
    fresh.edge_index, fresh.bond_mask = yb3.edge_index, yb3.bond_mask
    gout = prod.arch_module(fresh, torch.tensor([ctx.c_noise]), ctx.r_cut).pos
    want = (xn.pos - ctx.c_skip * yb2.pos) / ctx.c_out
    assert torch.allclose(gout, want, atol=2e-5), (gout - want).abs().max()


def test_sampler_end_to_end_with_callbacks(models):
    """jamun.sampling.Sampler.sample: outer loop over batches, chain continuation, callback protocol, unbatching."""
    from jamun_b200 import data
    from jamun_b200.sampling import Sampler
    from jamun_b200.sampling.mcmc import BAOAB
    from jamun_b200.sampling.walkjump import SingleMeasurementSampler

    o32, o64, prod, t, y = _setup(models, [22, 15, 9], seed=2)

    class Rec:
        def __init__(self):
            self.events, self.samples = [], []

        def on_sample_start(self, sampler):
            self.events.append("start")

        def on_before_sample_batch(self, sampler):
            self.events.append("before")

        def on_after_sample_batch(self, sample, sampler):
            self.events.append("after")
            self.samples.append(sample)

        def on_sample_end(self, sampler):
            self.events.append("end")

    rec = Rec()
    torch.manual_seed(5)
    sampler = Sampler(callbacks=[rec])
    bs = SingleMeasurementSampler(BAOAB(delta=0.04, friction=1.0, M=1.0, steps=5, save_trajectory=True, save_every_n_steps=2,
                                        score_fn_clip=100.0), SIGMA)
    sampler.sample(prod, bs, num_batches=2, init_graphs=data.Batch.from_tensors(t), continue_chain=True)
    assert rec.events == ["start", "before", "after", "before", "after", "end"]
    graphs = rec.samples[0]
    assert len(graphs) == 3 and graphs[0]["xhat_traj"].shape == (22, 3, 3) and graphs[2]["sample"].shape == (9, 3)
    assert all(torch.isfinite(gph["xhat"]).all() for gph in graphs)


def test_full_size_properties(models, monkeypatch):
    """BASELINE.json's headline size (1024 uncapped 2AA chains, 18 k atoms: every kernel runs multi-wave / persistent with
    all 148 SMs busy) through size-independent properties: (1) the tensor-core pipeline equals the exact-fp32 CUDA-core
    pipeline; (2) a chain's result does not depend on which other chains share the batch (oracle on a 12-chain slice,
    bit-level neighbour lists included); (3) rotating + translating the input rotates the output."""
    from jamun_b200 import data, engine, synthetic
    from oracle import jamun_oracle as O

    o32, o64, prod = models
    sizes = synthetic.workload_sizes("2AA", 1024)
    t = synthetic.make_tensors(sizes)
    gen = torch.Generator().manual_seed(17)
    y = t["pos"] + SIGMA * torch.randn(t["pos"].shape, generator=gen)
    batch = data.Batch.from_tensors(t).to("cuda")

    def run(pos):
        yb = batch.clone("pos")
        yb.pos = pos.cuda().contiguous()
        return prod.xhat(yb, SIGMA).pos.cpu()

    x = run(y)
    assert torch.isfinite(x).all()
    # (1) exact-fp32 conv + block-tail kernels at the same size
    monkeypatch.setattr(engine, "CONV_IMPL", "simt")
    monkeypatch.setenv("JAMUN_B200_TAIL", "simt")
    x_simt = run(y)
    monkeypatch.setattr(engine, "CONV_IMPL", "tc")
    monkeypatch.delenv("JAMUN_B200_TAIL")
    assert torch.allclose(x, x_simt, rtol=1e-4, atol=1e-5), (x - x_simt).abs().max()
    # (1b) the tf32-split GEMMs with the unfused block tail (round-1 pipeline) agree with the fp16-split, fused-epilogue default
    monkeypatch.setenv("JAMUN_B200_GEMM", "tf32")
    x_tf32 = run(y)
    monkeypatch.delenv("JAMUN_B200_GEMM")
    assert torch.allclose(x, x_tf32, rtol=1e-4, atol=1e-5), (x - x_tf32).abs().max()
    monkeypatch.setenv("JAMUN_B200_TAIL_FUSE", "0")
    x_unfused = run(y)
    monkeypatch.delenv("JAMUN_B200_TAIL_FUSE")
    assert torch.allclose(x, x_unfused, rtol=1e-5, atol=1e-6), (x - x_unfused).abs().max()
    # (1c) the operand workspace capped so that the batch is processed in two row chunks (as BASELINE config 4 is at 512 k
    # atoms): fused epilogues address the block-tail operands and the path-2 addend by chunk offset
    monkeypatch.setattr(engine.Topology, "WORKSPACE_BYTES", 9600 * 65 * 11 * 32 * 4)
    batch_c = data.Batch.from_tensors(t).to("cuda")
    yb = batch_c.clone("pos")
    yb.pos = y.cuda().contiguous()
    x_chunked = prod.xhat(yb, SIGMA).pos.cpu()
    assert prod.topology_for(yb).chunk_rows == 9600
    monkeypatch.undo()
    # (not bit-identical: the second chunk of the initial block has few enough row tiles to take the split-K contraction)
    assert torch.allclose(x_chunked, x, rtol=1e-5, atol=1e-6), (x_chunked - x).abs().max()
    # (2) the last 12 chains alone, against the oracle
    k = 12
    n_tail = sum(sizes[-k:])
    t_tail = synthetic.make_tensors(sizes[-k:], first_chain_id=len(sizes) - k)
    assert torch.equal(t_tail["pos"], t["pos"][-n_tail:])
    with torch.no_grad():
        ref = o32.xhat(make_oracle_batch(t_tail).with_pos(y[-n_tail:]), SIGMA)
    assert torch.allclose(x[-n_tail:], ref, rtol=1e-4, atol=1e-5), (x[-n_tail:] - ref).abs().max()
    # (3) equivariance
    Q, _ = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(1)))
    if torch.det(Q) < 0:
        Q[:, 0] *= -1
    xr = run(y @ Q.T + torch.tensor([0.5, -0.1, 0.2]))
    assert torch.allclose(xr, x @ Q.T, atol=3e-5), (xr - x @ Q.T).abs().max()


def test_dense_graph_tc_equals_exact_pipeline(models, monkeypatch):
    """1000-atom chains: every in-degree sits at the neighbour cap (33-35 > 32, so each node takes two operand chunks in the
    tensor-core builder); the tensor-core pipeline must still equal the exact-fp32 CUDA-core pipeline."""
    from jamun_b200 import data, engine, synthetic

    o32, o64, prod = models
    sizes = [1000] * 5 + [333, 150]
    t = synthetic.make_tensors(sizes, n_res=100)
    gen = torch.Generator().manual_seed(23)
    y = t["pos"] + SIGMA * torch.randn(t["pos"].shape, generator=gen)
    yb = data.Batch.from_tensors(t).to("cuda")
    yb.pos = y.cuda().contiguous()

    def run():
        return prod.xhat(yb, SIGMA).pos.cpu()

    x = run()
    topo = prod.topology_for(yb)
    deg = (topo.rowptr[1:] - topo.rowptr[:-1])
    assert int(deg.max()) > 32 and float(deg.float().mean()) > 30
    monkeypatch.setattr(engine, "CONV_IMPL", "simt")
    monkeypatch.setenv("JAMUN_B200_TAIL", "simt")
    x_simt = run()
    assert torch.isfinite(x).all()
    assert torch.allclose(x, x_simt, rtol=1e-4, atol=1e-5), (x - x_simt).abs().max()
    # the last (150-atom, in-degree ~ cap) chain alone against the oracle: chains are independent
    n_tail = sizes[-1]
    t_tail = synthetic.make_tensors(sizes[-1:], n_res=100, first_chain_id=len(sizes) - 1)
    assert torch.equal(t_tail["pos"], t["pos"][-n_tail:])
    with torch.no_grad():
        ref = o32.xhat(make_oracle_batch(t_tail).with_pos(y[-n_tail:]), SIGMA)
    assert torch.allclose(x[-n_tail:], ref, rtol=1e-4, atol=1e-5), (x[-n_tail:] - ref).abs().max()


def test_4aa_size_oracle_slice(models):
    """BASELINE config 3's per-GPU shape (1024 uncapped 4AA chains, ~35 k atoms, in-degree ~ 20): the last 8 chains against the
    oracle at north_star's tolerance (xhat rtol 1e-4 / atol 1e-5; scores the same relative error)."""
    from jamun_b200 import data, synthetic

    o32, o64, prod = models
    sizes = synthetic.workload_sizes("4AA", 1024)
    t = synthetic.make_tensors(sizes, n_res=4)
    y = t["pos"] + SIGMA * torch.randn(t["pos"].shape, generator=torch.Generator().manual_seed(29))
    yb = data.Batch.from_tensors(t).to("cuda")
    yb.pos = y.cuda().contiguous()
    x = prod.xhat(yb, SIGMA).pos.cpu()
    sc = prod.score(yb, SIGMA).cpu()
    assert torch.isfinite(x).all()
    k = 8
    n_tail = sum(sizes[-k:])
    t_tail = synthetic.make_tensors(sizes[-k:], n_res=4, first_chain_id=len(sizes) - k)
    assert torch.equal(t_tail["pos"], t["pos"][-n_tail:])
    with torch.no_grad():
        ob = make_oracle_batch(t_tail).with_pos(y[-n_tail:])
        ref, sref = o32.xhat(ob, SIGMA), o32.score(ob, SIGMA)
    assert torch.allclose(x[-n_tail:], ref, rtol=1e-4, atol=1e-5), (x[-n_tail:] - ref).abs().max()
    assert ((sc[-n_tail:] - sref).norm() / sref.norm()).item() <= 1e-4


def test_full_size_graph_replay_equals_eager(models):
    """The same check at the headline size (1024 chains): with two streams and persistent kernels in the captured step, graph
    replay must reproduce the eager launches bit for bit, and a second run of the same seed must reproduce the first."""
    import itertools

    from jamun_b200 import data, synthetic
    from jamun_b200.sampling.mcmc.functional import _splitting, fused_baoab

    o32, o64, prod = models
    t = synthetic.make_tensors(synthetic.workload_sizes("2AA", 1024))
    y = (t["pos"] + SIGMA * torch.randn(t["pos"].shape, generator=torch.Generator().manual_seed(4))).cuda()
    topo = prod.topology_for(data.Batch.from_tensors(t).to("cuda"))
    kw = dict(steps=6, v_init="gaussian", save_trajectory=False, delta=0.04, friction=1.0, M=1.0, inverse_temperature=1.0,
              score_fn_clip=100.0)
    outs = []
    for use_graph in (False, True, True):
        torch.manual_seed(7)
        _splitting._call_counter = itertools.count(1)
        outs.append(fused_baoab(prod, topo, y, SIGMA, use_cuda_graph=use_graph, **kw))
    for key in ("y", "v", "xhat"):
        assert torch.isfinite(outs[0][key]).all()
        assert torch.equal(outs[0][key], outs[1][key]), key
        assert torch.equal(outs[1][key], outs[2][key]), key
