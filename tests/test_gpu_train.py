"""Row a16: training forward + backward on the library's kernels (torch.library operators of jamun_b200/autograd_ops.py).

Every operator is checked against torch autograd over a plain-torch evaluation of the same math (tests/torch_reference.py, the
round-1 library path, now test infrastructure); the whole network's parameter gradients are checked against the oracle's
autograd (reference formulation, CPU)."""
import math

import pytest
import torch

import torch_reference as R
from conftest import make_oracle_batch

pytestmark = pytest.mark.gpu
SIGMA = 0.04


def _setup(models, sizes, seed=3):
    from jamun_b200 import data, synthetic

    o32, o64, prod = models
    t = synthetic.make_tensors(sizes)
    gen = torch.Generator().manual_seed(seed)
    y = t["pos"] + SIGMA * torch.randn(t["pos"].shape, generator=gen)
    batch = data.Batch.from_tensors(t).to("cuda")
    yb = batch.clone("pos")
    yb.pos = y.cuda()
    return o32, prod, t, y, batch, yb


def _graph(prod, yb):
    """Topology with a CSR + edge geometry for yb, and the block operands (leaf copies requiring grad)."""
    from jamun_b200 import ops

    topo = prod.topology_for(yb)
    ctx = prod.sigma_context(SIGMA)
    ybar, p = ops.center_scale(yb.pos.contiguous(), topo.chain_ptr, ctx.c_in)
    topo.build_csr(ybar, ctx.r_cut)
    mu, step = prod.arch_module.plan(ctx.c_noise, "cuda").radial_grid(ctx.r_cut)
    ops.edge_geom(p, topo.rowptr, topo.col, topo.edst, mu, step, topo.rhat, topo.rb)
    torch.cuda.synchronize()
    return topo, ctx, p


def _leaf(t):
    return t.detach().clone().requires_grad_(True)


def _close(got, want, tol, what):
    scale = max(want.abs().max().item(), 1e-12)
    err = (got - want).abs().max().item()
    assert err <= tol * scale + 1e-9, f"{what}: err {err:.3e} scale {scale:.3e}"


@pytest.mark.parametrize("layer,sizes", [(0, [22, 15, 9, 30, 1, 2]), (1, [22, 15, 9, 30, 1, 2]), (1, [150, 40])])
def test_conv_operator_forward_and_backward_vs_torch(models, layer, sizes):
    """[150, 40]: a dense graph -- in-degrees beyond 32 (neighbour cap + bonded edges), so the per-edge backward takes a second
    batch per receiver and sources have more than 32 out-edges."""
    import jamun_b200.autograd_ops  # noqa: F401

    o32, prod, t, y, batch, yb = _setup(models, sizes)
    topo, ctx, p = _graph(prod, yb)
    if len(sizes) == 2:
        deg = (topo.rowptr[1:] - topo.rowptr[:-1]).max().item()
        outdeg = (topo.src_rowptr[1:] - topo.src_rowptr[:-1]).max().item()
        assert deg > 32 and outdeg > 32, (deg, outdeg)
    g = prod.arch_module
    blk = [g.initial_projector, *g.layers][layer]
    pk = blk.pack(g.embed_bondedness.weight)
    s_in, v_in = pk["s_in"], pk["v_in"]
    E, N = int(topo.rowptr[-1]), topo.N
    gen = torch.Generator().manual_seed(layer)
    x = torch.randn(N, s_in + 3 * v_in, generator=gen).cuda()
    h = torch.zeros(topo.cap, 64, device="cuda")
    h[:E] = torch.randn(E, 64, generator=gen).cuda() * 0.5
    dout = torch.randn(N, 248, generator=gen).cuda()
    xa, ha, m0a, m1a = _leaf(x), _leaf(h), _leaf(pk["m0"]), _leaf(pk["m1"])
    out = torch.ops.jamun_b200.conv(xa, ha, topo.rhat, topo.rowptr, topo.col, topo.edst, topo.src_rowptr, topo.src_eid, m0a, m1a,
                                    s_in, v_in, pk["alpha0"], pk["alpha1"])
    out.backward(dout)
    # torch reference (autograd)
    xb, hb, m0b, m1b = _leaf(x), _leaf(h[:E]), _leaf(pk["m0"]), _leaf(pk["m1"])
    eid_pad, deg = R._padded_edge_index(topo.rowptr, E)
    ref = R._conv(xb, s_in, v_in, topo.col[:E].long(), eid_pad, deg, hb, topo.rhat[:E, :3], dict(pk, m0=m0b, m1=m1b))
    ref.backward(dout)
    _close(out.detach(), ref.detach(), 1e-4, "conv out")
    _close(xa.grad, xb.grad, 2e-4, "dx")
    _close(ha.grad[:E], hb.grad, 2e-4, "dh")
    _close(m0a.grad, m0b.grad, 2e-4, "dm0")
    _close(m1a.grad, m1b.grad, 2e-4, "dm1")


@pytest.mark.parametrize("layer", [0, 1, 5])
def test_block_tail_operator_vs_torch(models, layer):
    import jamun_b200.autograd_ops  # noqa: F401

    o32, prod, t, y, batch, yb = _setup(models, [22, 15, 9, 30])
    g = prod.arch_module
    blk = [g.initial_projector, *g.layers][layer]
    pk = blk.pack(g.embed_bondedness.weight)
    s_in, v_in = pk["s_in"], pk["v_in"]
    N = t["pos"].shape[0]
    gen = torch.Generator().manual_seed(10 + layer)
    rnd = lambda *s: torch.randn(*s, generator=gen).cuda()  # noqa: E731
    conv_out, x_in = rnd(N, 248), rnd(N, s_in + 3 * v_in)
    x_res = rnd(N, 216) if layer > 0 else None
    skip_w = torch.sigmoid(rnd(152)) if layer > 0 else None
    s_next = 1 + 0.1 * rnd(152) if layer < 5 else None
    names = ["wself_s", "wself_v", "wskip_s"] + (["wskip_v"] if v_in else [])

    def run(fn):
        leaves = dict(conv=_leaf(conv_out), x_in=_leaf(x_in), x_res=None if x_res is None else _leaf(x_res),
                      skip_w=None if skip_w is None else _leaf(skip_w), s_next=None if s_next is None else _leaf(s_next))
        leaves.update({k: _leaf(pk[k]) for k in names})
        x_new, x_scaled = fn(leaves)
        loss = (x_new * d_new).sum() + ((x_scaled * d_scaled).sum() if s_next is not None else 0.0)
        loss.backward()
        return x_new.detach(), (x_scaled.detach() if s_next is not None else None), leaves

    d_new, d_scaled = rnd(N, 216), rnd(N, 216)
    ours = run(lambda L: torch.ops.jamun_b200.block_tail(L["conv"], L["x_in"], L["x_res"], L["wself_s"], L["wself_v"], L["wskip_s"],
                                                         L.get("wskip_v"), L["skip_w"], L["s_next"], s_in, v_in, pk["c_act"],
                                                         pk["c_gate"]))
    ref = run(lambda L: R._block_tail(L["conv"], L["x_in"], s_in, v_in, L["x_res"],
                                      dict(pk, **{k: L[k] for k in names}), L["skip_w"], L["s_next"]))
    _close(ours[0], ref[0], 1e-5, "x_new")
    if s_next is not None:
        _close(ours[1], ref[1], 1e-5, "x_scaled")
    for k, leaf in ours[2].items():
        if leaf is not None:
            _close(leaf.grad, ref[2][k].grad, 1e-4, f"d{k}")


def test_small_operators_vs_torch(models):
    """head, radial_hidden, atom_embed, noise_mlp, combine_xhat, coordinate_loss: forward and every gradient."""
    import torch.nn.functional as F

    import jamun_b200.autograd_ops  # noqa: F401

    T = torch.ops.jamun_b200
    o32, prod, t, y, batch, yb = _setup(models, [22, 15, 9, 30, 1])
    topo, ctx, p = _graph(prod, yb)
    g = prod.arch_module
    N, E = topo.N, int(topo.rowptr[-1])
    gen = torch.Generator().manual_seed(99)
    rnd = lambda *s: torch.randn(*s, generator=gen).cuda()  # noqa: E731
    # head
    hb, lin2 = g.output_head[0], g.output_head[1]
    x, w1s, w1v, w2 = rnd(N, 216), hb.lin.packed(0).detach(), hb.lin.packed(1).detach(), rnd(32)
    dg = rnd(N, 3)
    la = [_leaf(v) for v in (x, w1s, w1v, w2)]
    T.head(*la, hb.gate.c_gate).backward(dg)
    lb = [_leaf(v) for v in (x, w1s, w1v, w2)]
    gate = hb.gate.c_gate * torch.sigmoid(lb[0][:, :120] @ lb[1][:, 120:])
    hv = lb[0][:, 120:].reshape(N, 3, 32) @ lb[2]
    ((hv * gate[:, None, :]) * lb[3]).sum(-1).backward(dg)
    for a, b, nm in zip(la, lb, ("x", "w1s", "w1v", "w2")):
        _close(a.grad, b.grad, 1e-4, f"head d{nm}")
    # radial hidden
    pk = g.layers[0].pack(g.embed_bondedness.weight)
    dh = torch.zeros(topo.cap, 64, device="cuda")
    dh[:E] = rnd(E, 64)
    la = [_leaf(pk["w0r"]), _leaf(pk["b0eff"])]
    h = T.radial_hidden(topo.rb, topo.ebond, topo.rowptr, *la)
    h.backward(dh)
    lb = [_leaf(pk["w0r"]), _leaf(pk["b0eff"])]
    z = topo.rb[:E] @ lb[0] + lb[1][topo.ebond[:E].long()]
    href = z * torch.sigmoid(z)
    href.backward(dh[:E])
    _close(h[:E].detach(), href.detach(), 1e-5, "radial h")
    _close(la[0].grad, lb[0].grad, 1e-4, "dw0r")
    _close(la[1].grad, lb[1].grad, 1e-4, "db0eff")
    # atom embedding
    tabs = [tb.detach() for tb in g.atom_embedder.tables()]
    scale = 1 + 0.1 * rnd(56)
    dx0 = rnd(N, 56)
    la = [_leaf(v) for v in (*tabs, scale)]
    T.atom_embed(*topo.idx, *la).backward(dx0)
    lb = [_leaf(v) for v in (*tabs, scale)]
    (torch.cat([tb[i.long()] for tb, i in zip(lb[:4], topo.idx)], dim=1) * lb[4]).backward(dx0)
    for k, (a, b) in enumerate(zip(la, lb)):
        _close(a.grad, b.grad, 1e-4, f"embed grad {k}")
    # noise MLP
    for mod, sig in ((g.noise_scalings[0], False), (g.skip_connections[0].weights, True), (g.initial_noise_scaling, False)):
        opsd = [o.detach() for o in mod.mlp_operands()]
        d = rnd(opsd[1].numel())
        la = [_leaf(v) for v in opsd]
        out = T.noise_mlp(*la, ctx.c_noise, sig)
        out.backward(d)
        lb = [_leaf(v) for v in opsd]
        ref = F.selu(lb[0] * ctx.c_noise + lb[1]) @ lb[2].T + lb[3]
        ref = torch.sigmoid(ref) if sig else ref
        ref.backward(d)
        _close(out.detach(), ref.detach(), 1e-5, "noise mlp")
        for k, (a, b) in enumerate(zip(la, lb)):
            _close(a.grad, b.grad, 1e-4, f"noise mlp grad {k}")
    # xhat combine + coordinate loss
    gnet, ybar, xt = rnd(N, 3), rnd(N, 3), rnd(N, 3)
    lw = torch.rand(topo.G, generator=gen).cuda() + 0.5
    dl = rnd(topo.G)
    ga = _leaf(gnet)
    xh = T.combine_xhat(ga, ybar, topo.chain_ptr, ctx.c_skip, ctx.c_out, True)
    loss, raw, rmsd = T.coordinate_loss(xh, xt, topo.chain_of, topo.chain_ptr, lw, 1.0 / ctx.c_out ** 2, SIGMA)
    loss.backward(dl)
    gb = _leaf(gnet)
    pre = ctx.c_skip * ybar + ctx.c_out * gb
    cnt = (topo.chain_ptr_long[1:] - topo.chain_ptr_long[:-1]).clamp_min(1).float()
    seg = lambda v: torch.zeros(topo.G, *v.shape[1:], device="cuda").index_add_(0, topo.batch_long, v) / cnt.reshape(-1, *([1] * (v.dim() - 1)))  # noqa: E731
    xr = pre - seg(pre)[topo.batch_long]
    rawr = seg(((xr - xt) ** 2).sum(-1))
    lossr = rawr * lw / ctx.c_out ** 2
    lossr.backward(dl)
    _close(xh.detach(), xr.detach(), 1e-5, "xhat")
    _close(loss.detach(), lossr.detach(), 1e-5, "loss")
    _close(raw, rawr.detach(), 1e-5, "raw")
    _close(rmsd, seg(((xr - xt) ** 2).sum(-1).sqrt()).detach() / (SIGMA * math.sqrt(3)), 1e-5, "rmsd")
    _close(ga.grad, gb.grad, 1e-4, "dg")


def test_kabsch_kernel_matches_oracle():
    from jamun_b200 import synthetic
    from jamun_b200.utils import kabsch_algorithm
    from oracle import jamun_oracle as O

    t = synthetic.make_tensors([22, 15, 9, 30, 2, 1, 57])
    gen = torch.Generator().manual_seed(4)
    x = t["pos"]
    # rotate + translate + perturb every chain differently
    y = x.clone()
    for c in range(t["num_graphs"]):
        m = t["batch"] == c
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen))
        if torch.det(q) < 0:
            q[:, 0] = -q[:, 0]
        y[m] = x[m] @ q.T + torch.randn(3, generator=gen) + 0.05 * torch.randn(int(m.sum()), 3, generator=gen)
    ref = O.kabsch_algorithm(y.double(), x.double(), t["batch"], t["num_graphs"]).float()
    got = kabsch_algorithm(y.cuda(), x.cuda(), t["batch"].cuda(), t["num_graphs"]).cpu()
    big = torch.bincount(t["batch"])[t["batch"]] >= 3  # for 1- and 2-atom chains the null-space part of R is arbitrary
    assert torch.allclose(got[big], ref[big], rtol=1e-4, atol=2e-5), (got[big] - ref[big]).abs().max()
    # ... but the aligned points themselves are determined
    assert torch.allclose(got[~big], ref[~big], rtol=1e-4, atol=2e-5), (got[~big] - ref[~big]).abs().max()


def test_gemm_wide_non_stationary_column_blocks():
    """5 K stages x many column blocks (the dA = G . M^T shape of the conv backward) through jamun_pack_b(transpose)."""
    from jamun_b200 import ops

    gen = torch.Generator().manual_seed(8)
    rows, K, Ncols = 300, 152, 1000
    rows_pad = 384
    A = torch.randn(rows, K, generator=gen).cuda()
    Mt = torch.randn(Ncols, K, generator=gen).cuda()  # weights stored [cols, K]: the GEMM needs the transposed image
    a_op = torch.empty(5 * rows_pad * 32, device="cuda")
    ops.pack_rows(A, 0, K, rows_pad, a_op)
    cb = (Ncols + 127) // 128
    img = ops.pack_b(Mt, n_stages=5, n_pad=128, k_src=Ncols, n_valid=K, col_blocks=cb, transpose=True)
    out = torch.full((rows, cb * 128), float("nan"), device="cuda")
    ops.gemm_tf32x3([a_op.data_ptr()], [img.data_ptr()], [5], [128], [128], [0], [1.0], rows, rows_pad, None, out.data_ptr(), cb * 128,
                    col_blocks=cb, b_block_floats=5 * 2 * 128 * 32)
    ref = A.double() @ Mt.double().T
    assert (out[:, :Ncols].double() - ref).abs().max() <= 1e-5 * ref.abs().max()
    assert float(out[:, Ncols:].abs().max()) == 0.0


def test_training_gradients_match_oracle_autograd(models):
    from oracle import jamun_oracle as O

    o32, prod, t, y, batch, yb = _setup(models, [22, 15, 9])
    ob = make_oracle_batch(t)
    o32.zero_grad()
    xh_o = o32.xhat(ob.with_pos(y), SIGMA)
    loss_o, _ = O.compute_loss(o32, ob, xh_o, SIGMA)
    loss_o.mean().backward()
    ref = {k: p.grad.clone() for k, p in o32.named_parameters() if p.grad is not None}
    o32.zero_grad()

    prod.train()
    prod.zero_grad()
    xh_p = prod.xhat_with_grad(yb, SIGMA)
    loss_p, aux = prod.compute_loss(batch, xh_p, SIGMA)
    loss_p.mean().backward()
    prod.eval()
    with torch.no_grad():
        assert torch.allclose(xh_p.pos.detach(), prod.xhat(yb, SIGMA).pos, rtol=1e-4, atol=1e-5)  # training fwd == sampling fwd
    assert torch.allclose(loss_p.detach().cpu(), loss_o.detach(), rtol=2e-4, atol=1e-6)
    got = {k.replace("g._orig_mod.", "g."): p.grad.detach().cpu() for k, p in prod.named_parameters() if p.grad is not None}
    prod.zero_grad()
    assert len(ref) > 40 and set(ref) <= set(got), sorted(set(ref) - set(got))[:5]
    worst = 0.0
    for k, g_ref in ref.items():
        scale = max(g_ref.abs().max().item(), 1e-12)
        err = (got[k] - g_ref).abs().max().item()
        worst = max(worst, err / scale)
        assert err <= 1e-3 * scale + 1e-7, f"{k}: grad err {err} scale {scale}"
    print(f"worst relative gradient error vs oracle autograd: {worst:.2e}")


def test_training_step_is_deterministic_and_reduces_loss(models, monkeypatch):
    """A few SGD steps through Denoiser.training_step on one fixed noisy batch lower the loss; two identical runs give
    bit-identical losses and gradients (no atomics anywhere on the backward path)."""
    import jamun_b200 as J
    from jamun_b200 import data, synthetic

    o32, _, prod0 = models
    t = synthetic.make_tensors([20, 20, 20, 20])

    def run():
        prod = J.default_denoiser()
        prod.load_state_dict(o32.state_dict())
        prod = prod.to("cuda").train()
        prod.add_fixed_noise = True  # same noise draw every step (denoiser.py:94-103)
        batch = data.Batch.from_tensors(t).to("cuda")
        opt = torch.optim.SGD(prod.parameters(), lr=1e-3)
        losses = []
        for it in range(4):
            opt.zero_grad()
            out = prod.training_step(batch, it)
            out["loss"].backward()
            opt.step()
            losses.append(float(out["loss"].detach()))
        return losses, [p.grad.clone() for p in prod.parameters() if p.grad is not None]

    l1, g1 = run()
    l2, g2 = run()
    assert all(torch.isfinite(torch.tensor(l1)))
    assert l1[-1] < l1[0], l1
    assert l1 == l2 and all(torch.equal(a, b) for a, b in zip(g1, g2))
    # operands kept from the forward (conv_keep, the default) vs rebuilt in the backward: the same bits
    monkeypatch.setenv("JAMUN_B200_TRAIN_KEEP_A_GB", "0")
    l3, g3 = run()
    assert l1 == l3 and all(torch.equal(a, b) for a, b in zip(g1, g3))


def test_operator_registrations_pass_opcheck(models):
    """torch.library.opcheck: schema, fake-tensor implementation and autograd registration of the operators."""
    import jamun_b200.autograd_ops  # noqa: F401

    o32, prod, t, y, batch, yb = _setup(models, [9, 12])
    topo, ctx, p = _graph(prod, yb)
    g = prod.arch_module
    N = topo.N
    gen = torch.Generator().manual_seed(5)
    hb = g.output_head[0]
    args = (torch.randn(N, 216, generator=gen).cuda().requires_grad_(True), hb.lin.packed(0).detach().requires_grad_(True),
            hb.lin.packed(1).detach().requires_grad_(True), torch.randn(32, generator=gen).cuda().requires_grad_(True), hb.gate.c_gate)
    torch.library.opcheck(torch.ops.jamun_b200.head.default, args, test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    pk = g.layers[0].pack(g.embed_bondedness.weight)
    x = torch.randn(N, 216, generator=gen).cuda().requires_grad_(True)
    h = torch.rand(topo.cap, 64, generator=gen).cuda().requires_grad_(True)
    args = (x, h, topo.rhat, topo.rowptr, topo.col, topo.edst, topo.src_rowptr, topo.src_eid, pk["m0"].detach().requires_grad_(True),
            pk["m1"].detach().requires_grad_(True), 120, 32, pk["alpha0"], pk["alpha1"])
    torch.library.opcheck(torch.ops.jamun_b200.conv.default, args, test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    torch.library.opcheck(torch.ops.jamun_b200.conv_keep.default, args, test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


@pytest.mark.parametrize("rows,n_stages,nslots,ncomp,W,mode", [(300, 10, 5, 1, 152, 0), (1000, 6, 2, 3, 32, 0), (77, 65, 1, 1, 120, 1),
                                                              (4000, 8, 2, 1, 56, 0)])
def test_stage_atb_tensor_core_matches_simt_and_fp64(rows, n_stages, nslots, ncomp, W, mode):
    """dW = A^T . B over the nodes: tcgen05 kernel (MN-major operands, TS mode, split over the node range) vs the CUDA-core
    kernel vs fp64."""
    from jamun_b200 import ops

    gen = torch.Generator().manual_seed(rows + W)
    rows_pad = (rows + 127) // 128 * 128
    A = torch.randn(ncomp, n_stages * 32, rows, generator=gen)          # [c][(stage, u)][row]
    B = torch.randn(rows, 40 + ncomp * W, generator=gen)
    a_op = torch.full((ncomp, n_stages, rows_pad, 32), float("nan"))   # stale rows beyond `rows` must be ignored
    a_op[:, :, :rows] = A.reshape(ncomp, n_stages, 32, rows).permute(0, 1, 3, 2)
    rr = torch.arange(rows_pad)[:, None]
    ll = torch.arange(32)[None, :]
    pos = (((ll // 4) ^ (rr % 8)) * 4 + ll % 4).expand(ncomp, n_stages, rows_pad, 32)
    a_sw = torch.empty_like(a_op).scatter_(3, pos, a_op).cuda().contiguous()
    Bc = B.cuda().contiguous()
    ref = sum(A[c].double() @ B[:, 40 + c * W: 40 + (c + 1) * W].double() for c in range(ncomp))  # [(stage,u), W]
    K = n_stages // nslots
    if mode == 0:
        slot_row0, slot_rows = [40 * s for s in range(nslots)], [32 if s % 2 == 0 else 24 for s in range(nslots)]
        out_rows = 40 * nslots
    else:
        slot_row0 = slot_rows = None
        out_rows = W + 3
    outs = {}
    for impl in ("simt", "tc"):
        ops.ATB_IMPL = impl
        out = torch.zeros(K, out_rows, W if mode == 0 else 32, device="cuda")
        ops.stage_atb_auto(a_sw.data_ptr(), n_stages * rows_pad * 32, ncomp, n_stages, nslots, rows, rows_pad, Bc, 40, W, W, out, mode,
                           out_rows, slot_row0, slot_rows)
        torch.cuda.synchronize()
        outs[impl] = out.cpu()
    ops.ATB_IMPL = None
    r4 = ref.reshape(K, nslots, 32, W)
    want = torch.zeros_like(outs["tc"], dtype=torch.float64)
    if mode == 0:
        for s in range(nslots):
            want[:, slot_row0[s]:slot_row0[s] + slot_rows[s]] = r4[:, s, :slot_rows[s]]
    else:
        want[:, :W] = r4[:, 0].permute(0, 2, 1)
    scale = ref.abs().max().item()
    for impl in ("simt", "tc"):
        assert torch.isfinite(outs[impl]).all(), impl
        assert (outs[impl].double() - want).abs().max().item() <= 3e-5 * scale, impl


@pytest.mark.parametrize("rows,a,b,live", [(70_001, 32, 64, 66_003), (200, 32, 64, 200), (3000, 120, 120, 2500), (70_000, 20, 40, 70_000)])
def test_rowmat_dw_shapes_vs_fp64(rows, a, b, live):
    """dW = X^T . dY: the register-tiled CUDA-core kernel and, for narrow shapes with many rows (the radial MLP over the edges),
    the tensor-core kernel -- device-side row count, column offsets, ragged tails, bit-reproducible."""
    from jamun_b200 import ops

    gen = torch.Generator().manual_seed(rows + a)
    X = torch.randn(rows, a + 5, generator=gen).cuda()
    dY = torch.randn(rows, b + 3, generator=gen).cuda()
    rows_dev = torch.tensor([live], dtype=torch.int32, device="cuda")
    dW = torch.empty(a, b, device="cuda")
    ops.rowmat_dw(X, 2, dY, 1, dW, 0, a, b, rows=rows, rows_dev=rows_dev)
    want = X[:live, 2:2 + a].double().T @ dY[:live, 1:1 + b].double()
    scale = want.abs().max().item()
    assert (dW.double() - want).abs().max().item() <= 2e-5 * scale
    dW2 = torch.empty(a, b, device="cuda")
    ops.rowmat_dw(X, 2, dY, 1, dW2, 0, a, b, rows=rows, rows_dev=rows_dev)
    assert torch.equal(dW, dW2)


@pytest.mark.parametrize("rows,a,b,trans,acc,live", [(5000, 120, 120, False, False, 5000), (4101, 56, 120, True, True, 3999),
                                                      (2048, 32, 32, False, True, 2048), (9000, 152, 160, True, False, 9000),
                                                      (300, 120, 120, False, False, 300), (70_000, 64, 64, False, False, 66_000)])
def test_rowmat_mul_shapes_vs_fp64(rows, a, b, trans, acc, live):
    """Y (+)= X . W: the tensor-core kernel (>= 2048 rows) and the CUDA-core kernel -- transposed weights, accumulation,
    column offsets, ragged rows / columns, device-side row count; rows beyond the live count untouched."""
    from jamun_b200 import ops

    gen = torch.Generator().manual_seed(rows + a + b)
    X = torch.randn(rows, a + 3, generator=gen).cuda()
    W = (torch.randn(b, a + 2, generator=gen) if trans else torch.randn(a, b + 2, generator=gen)).cuda()
    Y0 = torch.randn(rows, b + 4, generator=gen).cuda()
    Y = Y0.clone()
    rows_dev = torch.tensor([live], dtype=torch.int32, device="cuda")
    ops.rowmat_mul(X, 1, W, 2 if trans else 1, Y, 3, a, b, trans_w=trans, accumulate=acc, rows=rows, rows_dev=rows_dev)
    Wm = W[:, 2:2 + a].T if trans else W[:, 1:1 + b]
    want = X[:live, 1:1 + a].double() @ Wm.double() + (Y0[:live, 3:3 + b].double() if acc else 0.0)
    scale = want.abs().max().item()
    assert (Y[:live, 3:3 + b].double() - want).abs().max().item() <= 2e-5 * scale
    assert torch.equal(Y[live:], Y0[live:]) and torch.equal(Y[:, :3], Y0[:, :3]) and torch.equal(Y[:, 3 + b:], Y0[:, 3 + b:])


def test_training_step_has_no_host_synchronisation(models):
    """After the first steps (Topology built, caches warm) a training step with sigma drawn from a continuous distribution --
    a new cut-off, radial grid and loss weight every step -- issues no synchronising CUDA call: the host runs ahead of the GPU."""
    import jamun_b200 as J
    from jamun_b200 import data, distributions, synthetic

    o32, _, _ = models
    prod = J.default_denoiser(sigma_distribution=distributions.UniformSigma(0.08, 0.02))
    prod.load_state_dict(o32.state_dict())
    prod = prod.to("cuda").train()
    batch = data.Batch.from_tensors(synthetic.make_tensors([20, 17, 9, 30])).to("cuda")
    opt = torch.optim.Adam(prod.parameters(), lr=1e-4)
    torch.manual_seed(3)

    def step():
        opt.zero_grad(set_to_none=True)
        out = prod.training_step(batch, 0)
        out["loss"].backward()
        opt.step()
        return out

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")
    try:
        outs = [step() for _ in range(3)]
    finally:
        torch.cuda.set_sync_debug_mode("default")
    sig = [float(o["sigma"]) for o in outs]
    assert len(set(sig)) == 3 and all(0.02 <= s <= 0.08 for s in sig)
    assert all(torch.isfinite(o["loss"]).item() for o in outs)


def test_atom_embedding_backward_with_many_atoms(models):
    """The table gradients with the atoms split over several CTAs per table row (N > 1024: per-split partial rows summed in
    order) against torch, and bit-reproducible."""
    import jamun_b200.autograd_ops  # noqa: F401

    T = torch.ops.jamun_b200
    _, _, prod = models
    g = prod.arch_module
    gen = torch.Generator().manual_seed(17)
    N = 5000 + 37
    idx = [torch.randint(0, hi, (N,), generator=gen, dtype=torch.int32).cuda() for hi in (5, 7, 21, 1)]
    tabs = [tb.detach() for tb in g.atom_embedder.tables()]
    scale = (1 + 0.1 * torch.randn(56, generator=gen)).cuda()
    dx0 = torch.randn(N, 56, generator=gen).cuda()
    grads = []
    for _ in range(2):
        la = [_leaf(v) for v in (*tabs, scale)]
        T.atom_embed(*idx, *la).backward(dx0)
        grads.append([a.grad.clone() for a in la])
    lb = [_leaf(v) for v in (*tabs, scale)]
    (torch.cat([tb[i.long()] for tb, i in zip(lb[:4], idx)], dim=1) * lb[4]).backward(dx0)
    for k, (a, b) in enumerate(zip(grads[0], lb)):
        _close(a, b.grad, 1e-4, f"embed grad {k}")
    assert all(torch.equal(a, b) for a, b in zip(*grads))
