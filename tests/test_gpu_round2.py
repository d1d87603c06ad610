"""GPU tests added in round 2: device-side weight packing, deterministic out-edge lists, lifetime of the tensors a cached
CUDA graph points at, the FullyConnectedTensorProduct seam, stale-CSR detection, plan-building launch count."""
import itertools
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu
SIGMA = 0.04


def _setup(models, sizes, seed=0):
    from jamun_b200 import synthetic

    o32, o64, prod = models
    t = synthetic.make_tensors(sizes)
    gen = torch.Generator().manual_seed(seed)
    y = t["pos"] + SIGMA * torch.randn(t["pos"].shape, generator=gen)
    return o32, o64, prod, t, y


@pytest.mark.parametrize("s_in,v_in", [(120, 32), (56, 0)])
def test_pack_b_kernel_equals_host_packing_bit_for_bit(s_in, v_in):
    from jamun_b200 import ops, packing

    gen = torch.Generator().manual_seed(s_in)
    m0 = torch.randn(65, s_in + v_in, 152, generator=gen).cuda()
    m1 = torch.randn(65, s_in + 2 * v_in, 32, generator=gen).cuda()
    b0, b1, wy = packing.pack_conv_operands_device(m0, m1, s_in, v_in)
    w0, w1, wyh = packing.conv_k_layout(m0, m1, s_in, v_in)
    assert torch.equal(b0, packing.pack_b_images(w0, 160))
    if v_in:
        assert torch.equal(b1, packing.pack_b_images(w1, 32))
    else:
        assert b1 is None
    wy_pad = torch.zeros(wyh.shape[0], 17 * 128, device="cuda")
    wy_pad[:, :wyh.shape[1]] = wyh
    assert torch.equal(wy, packing.pack_b_column_blocks(wy_pad, 128))
    # plain matrix, K not a multiple of 32 and N < n_pad: zero padding
    w = torch.randn(70, 20, generator=gen).cuda()
    wp = torch.zeros(96, 20, device="cuda")
    wp[:70] = w
    assert torch.equal(ops.pack_b(w, n_stages=3, n_pad=32), packing.pack_b_images(wp, 32))


def test_csr_by_source_lists_are_sorted(models):
    from jamun_b200 import data

    o32, o64, prod, t, y = _setup(models, [57, 3, 1, 40, 40, 17], seed=2)
    batch = data.Batch.from_tensors(t).to("cuda")
    topo = prod.topology_for(batch)
    topo.build_csr(y.cuda().contiguous(), 0.5872643)
    E = int(topo.rowptr[-1])
    rp, eid, col = topo.src_rowptr.cpu(), topo.src_eid.cpu(), topo.col.cpu()
    assert int(rp[-1]) == E and sorted(eid[:E].tolist()) == list(range(E))
    for j in range(topo.N):
        seg = eid[rp[j]:rp[j + 1]]
        assert torch.equal(seg, torch.sort(seg).values)
        assert bool((col[seg.long()] == j).all())


def test_cached_graph_survives_intervening_allocations(models):
    """the captured graph holds raw pointers (radial-basis centres); small allocations made and retained
    between two graph-mode calls must not change what a replay reads."""
    from jamun_b200 import data
    from jamun_b200.sampling.mcmc.functional import _splitting, fused_baoab

    o32, o64, prod, t, y = _setup(models, [22, 15, 9, 30, 12], seed=11)
    batch = data.Batch.from_tensors(t).to("cuda")
    topo = prod.topology_for(batch)
    kw = dict(steps=8, v_init="gaussian", save_trajectory=True, delta=0.04, friction=1.0, M=1.0, inverse_temperature=1.0,
              score_fn_clip=100.0)
    outs, keep = [], []
    for use_graph in (False, True, True, True):
        torch.manual_seed(321)
        _splitting._call_counter = itertools.count(1)
        outs.append(fused_baoab(prod, topo, y.cuda(), SIGMA, use_cuda_graph=use_graph, **kw))
        torch.cuda.synchronize()
        # grab (and keep) every small block the caching allocator could hand back, filled with garbage
        keep.append([torch.full((n,), 1e30, device="cuda") for n in (8, 32, 64, 128, 128, 256, 512) for _ in range(16)])
    for o in outs[1:]:
        for key in ("y", "xhat", "y_traj", "xhat_traj", "score_traj"):
            assert torch.equal(outs[0][key], o[key]), key


def test_fully_connected_tensor_product_forward_vs_oracle():
    from jamun_b200.e3tools.nn import FullyConnectedTensorProduct
    from oracle import jamun_oracle as O

    gen = torch.Generator().manual_seed(5)
    for in1 in ("120x0e+32x1e", "8x0e+8x0e+32x0e+8x0e", "2x0e+1x1e"):
        out = "152x0e+32x1e" if in1 != "2x0e+1x1e" else "1x0e+1x1e"
        tp = FullyConnectedTensorProduct(in1, "1x0e+1x1e", out)
        ref = O.FullyConnectedTP(in1, "1x0e+1x1e", out)
        assert tp.weight_numel == ref.weight_numel
        Z = 37
        x = torch.randn(Z, tp.irreps_in1.dim, generator=gen)
        sh = torch.randn(Z, 4, generator=gen)
        w = torch.randn(Z, tp.weight_numel, generator=gen)
        got = tp(x.cuda(), sh.cuda(), w.cuda()).cpu()
        want = ref(x.double(), sh.double(), w.double()).float()
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * want.abs().max()), (got - want).abs().max()


def test_stale_csr_tag_is_detected(models):
    """a graph whose positions changed after add_edges (or whose topology was rebuilt by another call) must not be
    evaluated on the old CSR."""
    from jamun_b200 import data

    o32, o64, prod, t, y = _setup(models, [22, 15, 9], seed=4)
    batch = data.Batch.from_tensors(t).to("cuda")
    yb = batch.clone("pos")
    yb.pos = y.cuda()
    ctx = prod.sigma_context(SIGMA)
    ref = prod.xhat_normalized(yb, SIGMA).pos.clone()
    yb2 = prod.add_edges(yb, ctx.r_cut)
    y_scaled = yb2.clone("pos")
    y_scaled.pos = yb2.pos * ctx.c_in  # tag of yb2 no longer describes these positions
    with pytest.raises(RuntimeError, match="current CSR"):
        prod.g(y_scaled, torch.tensor([ctx.c_noise]), ctx.r_cut)
    # another evaluation on the shared topology in between: the materialised edge list is used instead of the stale CSR
    yb3 = prod.add_edges(yb, ctx.r_cut, materialize=True)
    other = batch.clone("pos")
    other.pos = (y + 0.3).cuda()
    prod.xhat(other, SIGMA)
    y_scaled = yb3.clone("pos")
    y_scaled.pos = yb3.pos * ctx.c_in
    g = prod.g(y_scaled, torch.tensor([ctx.c_noise]), ctx.r_cut).pos
    assert torch.allclose(ctx.c_skip * yb.pos + ctx.c_out * g, ref, rtol=1e-4, atol=1e-5)


def test_plan_build_is_a_few_dozen_library_launches(models):
    """building a plan must not bury the library's kernels under framework indexing launches."""
    from torch.profiler import ProfilerActivity, profile

    from jamun_b200 import engine, ops

    o32, o64, prod = models
    g = prod.arch_module
    n0 = ops.LAUNCHES
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        engine.E3ConvPlan(g, -0.8, "cuda")
        torch.cuda.synchronize()
    ours = ops.LAUNCHES - n0
    kernels = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
               and "memset" not in e.name.lower()]
    foreign = [e.name for e in kernels if "jamun" not in e.name and "jb::" not in e.name and "pack_b_kernel" not in e.name
               and "noise_mlp" not in e.name]
    assert ours <= 50, ours
    assert len(foreign) <= 10, foreign[:10]


def test_sampling_launchers_are_torch_library_operators(models):
    """the tensor front end is a thin torch.library layer over the C ABI -- schema, fake (meta) implementation and
    mutation annotations of a representative set of the sampling-path operators."""
    from jamun_b200 import ops

    T = torch.ops.jamun_b200
    for name in ("radius_csr", "radius_csr_cells", "csr_by_source", "edge_geom", "edge_radial_hidden_all", "baoab_step", "center_scale_",
                 "atom_embed_", "noise_mlp_", "head_", "block_tail_", "tail_mix", "pack_rows", "tensor_product", "layout_to_soa"):
        assert hasattr(T, name), name
    assert "Tensor(a" in str(T.center_scale_.default._schema)  # mutated outputs are declared
    # opcheck of a mutating launcher and of a functional one (schema + fake tensor)
    y = torch.randn(30, 3, device="cuda")
    ptr = torch.tensor([0, 12, 30], dtype=torch.int32, device="cuda")
    torch.library.opcheck(T.center_scale_.default, (y, ptr, 1.7, torch.empty_like(y), torch.empty_like(y), True),
                          test_utils=("test_schema", "test_faketensor"))
    x = torch.randn(7, 216, device="cuda")
    torch.library.opcheck(T.layout_to_soa.default, (x, 120, 32), test_utils=("test_schema", "test_faketensor"))
    # fake-tensor propagation through a functional operator: shapes without running a kernel
    from torch._subclasses.fake_tensor import FakeTensorMode

    with FakeTensorMode():
        w = torch.empty(9, 64, device="cuda")
        out = T.linear_act(torch.empty(5, 64, device="cuda"), w, torch.empty(9, device="cuda"), 1)
        assert out.shape == (5, 9)


def test_fp16_range_overflow_falls_back_to_tf32_kernels(monkeypatch):
    """Activations beyond the fp16 range (here: an atom embedding scaled by 3e5) raise the fp16-split GEMMs' status bit; the
    call is repeated on the tf32-split kernels with a warning, and equals a run that used them from the start."""
    import copy
    import warnings

    import jamun_b200 as J
    from jamun_b200 import data, synthetic

    monkeypatch.setenv("JAMUN_B200_GEMM", "f16")  # restored on teardown (the fallback rewrites the variable)
    torch.manual_seed(3)
    m = J.default_denoiser()
    with torch.no_grad():
        m.arch_module.output_gain.fill_(1.0)
        for p in m.arch_module.atom_embedder.parameters():
            p.mul_(3.0e5)
    m = m.cuda().eval()
    t = synthetic.make_tensors([22, 15, 9, 30])
    gen = torch.Generator().manual_seed(2)
    y = (t["pos"] + 0.04 * torch.randn(t["pos"].shape, generator=gen)).cuda()

    def run():
        yb = data.Batch.from_tensors(t).to("cuda")
        yb.pos = y.clone()
        return m.xhat(yb, 0.04).pos.cpu()

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        x = run()
    assert any("fp16 range" in str(i.message) for i in w), [str(i.message) for i in w]
    import os
    assert os.environ["JAMUN_B200_GEMM"] == "tf32"
    assert torch.isfinite(x).all()
    with warnings.catch_warnings(record=True) as w2:
        warnings.simplefilter("always")
        x_tf32 = run()
    assert not any("fp16 range" in str(i.message) for i in w2)
    assert torch.equal(x, x_tf32)


def test_radial_hidden_tensor_core_kernel_matches_fp64_and_cuda_core_kernel():
    """jamun_edge_radial_hidden_mma (mma.sync tf32, three-product split) against fp64 torch and the packed-FP32 kernel: all layers,
    ragged edge count (not a multiple of 16), capacity rows beyond the live edges untouched."""
    from jamun_b200 import ops

    gen = torch.Generator().manual_seed(3)
    E, cap, L = 1000 + 7, 1100, 6
    rb = torch.rand(cap, 32, generator=gen) * 0.9
    w = torch.randn(L, 32, 64, generator=gen) / 32 ** 0.5
    b = torch.randn(L, 2, 64, generator=gen) * 0.3
    flag = (torch.rand(cap, generator=gen) < 0.2).to(torch.uint8)
    rowptr = torch.tensor([0, 400, E], dtype=torch.int32)
    z = torch.einsum("ek,lko->leo", rb[:E].double(), w.double()) + b.double()[:, flag[:E].long()].reshape(L, E, 64)
    want = (z * torch.sigmoid(z)).float()
    dev = "cuda"
    h_mma = torch.full((L, cap, 64), -7.0, device=dev)
    h_ffma = torch.full((L, cap, 64), -7.0, device=dev)
    img = ops.radial_pack_frag(w.to(dev))
    ops.edge_radial_hidden_mma(rb.to(dev), flag.to(dev), rowptr.to(dev), img, b.to(dev).contiguous(), h_mma)
    ops.edge_radial_hidden_all(rb.to(dev), flag.to(dev), rowptr.to(dev), w.to(dev).contiguous(), b.to(dev).contiguous(), h_ffma)
    for got in (h_mma, h_ffma):
        assert torch.allclose(got[:, :E].cpu(), want, rtol=2e-6, atol=2e-6), (got[:, :E].cpu() - want).abs().max()
        assert torch.all(got[:, E:] == -7.0)
    assert (h_mma[:, :E] - h_ffma[:, :E]).abs().max().item() < 4e-6


def test_new_noise_level_reuses_the_packed_plan(models):
    """a new sigma (validation / training draw it from a continuous distribution) must not re-pack the weight images --
    only the 11 noise-conditioning MLP outputs are re-evaluated; results equal those of a freshly built plan, bit for bit."""
    from jamun_b200 import data, ops, synthetic

    _, _, prod = models
    batch = data.Batch.from_tensors(synthetic.make_tensors([22, 15, 9])).to("cuda")
    arch = prod.arch_module
    xa = prod.xhat(batch, 0.04).pos.clone()
    plan = arch._plan
    imgs = [b["b0_img"].data_ptr() for b in plan.blocks]
    n0 = ops.LAUNCHES
    xb = prod.xhat(batch, 0.07).pos.clone()
    launched = ops.LAUNCHES - n0
    assert arch._plan is plan and [b["b0_img"].data_ptr() for b in plan.blocks] == imgs
    assert launched < 11 + 80, launched  # 11 noise MLPs + one evaluation (~60 launches), no pack_b launches
    xa2 = prod.xhat(batch, 0.04).pos.clone()
    assert torch.equal(xa, xa2) and not torch.equal(xa, xb)
    arch._plan = None  # fresh plans at both levels
    assert torch.equal(prod.xhat(batch, 0.07).pos, xb)
    arch._plan = None
    assert torch.equal(prod.xhat(batch, 0.04).pos, xa)
