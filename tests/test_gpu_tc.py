"""GPU tests of the tensor-core path: the 3xTF32 tcgen05 GEMM against fp64 matmul, and the two-launch conv
(aggregate builder + GEMM) against the exact-fp32 SIMT kernel and the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_gemm(A_list, B_list, n_pads, rows, row_scale=None, alphas=None, kind="tf32"):
    from jamun_b200 import ops, packing

    rows_pad = (rows + 127) // 128 * 128
    dev = "cuda"
    a_bufs, b_bufs, n_stages, n_valid, out_col, scales = [], [], [], [], [], []
    col = 0
    for A, B, n_pad in zip(A_list, B_list, n_pads):
        K = A.shape[1]
        assert K % 32 == 0
        S = K // 32
        a_sm = torch.full((S, rows_pad, 32), float("nan"), device=dev)
        a_sm[:, :rows] = A.to(dev).reshape(rows, S, 32).permute(1, 0, 2)
        # the A operand's 16-byte chunks are XOR-swizzled with (row & 7)
        rr = torch.arange(rows_pad, device=dev)[:, None]
        ll = torch.arange(32, device=dev)[None, :]
        pos = (((ll // 4) ^ (rr % 8)) * 4 + ll % 4).expand(S, rows_pad, 32)
        a_sw = torch.empty_like(a_sm).scatter_(2, pos, a_sm)
        a_bufs.append(a_sw.contiguous())
        if kind == "f16":
            scales.append(ops.f16_scale(B))
            b_bufs.append(ops.pack_b_f16(B.to(dev).contiguous(), S, n_pad, scales[-1]))
        else:
            b_bufs.append(packing.pack_b_images(B.to(dev), n_pad))
        n_stages.append(S)
        n_valid.append(B.shape[1])
        out_col.append(col)
        col += B.shape[1]
    out = torch.full((rows, col), float("nan"), device=dev)
    alphas = alphas or [1.0] * len(A_list)
    if kind == "f16":
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        ops.gemm_f16x3([a.data_ptr() for a in a_bufs], [b.data_ptr() for b in b_bufs], n_stages, list(n_pads), n_valid, out_col,
                       [a / s for a, s in zip(alphas, scales)], rows, rows_pad, row_scale.data_ptr() if row_scale is not None else None,
                       out.data_ptr(), col, status=status)
        torch.cuda.synchronize()
        assert int(status.item()) == 0
        return out
    ops.gemm_tf32x3([a.data_ptr() for a in a_bufs], [b.data_ptr() for b in b_bufs], n_stages, list(n_pads), n_valid, out_col,
                    alphas, rows, rows_pad, row_scale.data_ptr() if row_scale is not None else None, out.data_ptr(), col)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("rows,K,N,n_pad", [(128, 32, 32, 32), (128, 64, 152, 160), (300, 32 * 9, 152, 160), (77, 32 * 13, 32, 32),
                                            (1000, 32 * 40, 16, 16)])
@pytest.mark.parametrize("kind", ["tf32", "f16"])
def test_gemm_tf32x3_matches_fp64(rows, K, N, n_pad, kind):
    gen = torch.Generator().manual_seed(rows + K + N)
    A = torch.randn(rows, K, generator=gen)
    B = torch.randn(K, N, generator=gen) * 0.3
    out = _run_gemm([A], [B], [n_pad], rows, kind=kind).cpu().double()
    ref = A.double() @ B.double()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert torch.isfinite(out).all()
    assert err <= 3e-5 * scale, f"err {err} scale {scale}"  # tensor-core fp32 accumulation truncates (not RN)
    # a single TF32 pass would be ~1e-3 relative: make sure the split is really active
    assert err <= 1e-4 * (A.abs().double() @ B.abs().double()).max().item() / 30


@pytest.mark.parametrize("kind", ["tf32", "f16"])
def test_gemm_multi_segment_scaling_and_wraparound(kind):
    gen = torch.Generator().manual_seed(7)
    rows = 260
    A0, A1, A2, A3 = (torch.randn(rows, 32 * s, generator=gen) for s in (7, 5, 5, 5))
    B0 = torch.randn(32 * 7, 152, generator=gen)
    B1 = torch.randn(32 * 5, 32, generator=gen)
    rs = torch.rand(rows, generator=gen) + 0.5
    out = _run_gemm([A0, A1, A2, A3], [B0, B1, B1, B1], [160, 32, 32, 32], rows, row_scale=rs.cuda(), alphas=[0.5, 2.0, 2.0, 2.0], kind=kind)
    ref = torch.cat([0.5 * A0.double() @ B0.double()] + [2.0 * a.double() @ B1.double() for a in (A1, A2, A3)], dim=1) * rs.double()[:, None]
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= 3e-6 * ref.abs().max().item(), err


@pytest.mark.parametrize("kind", ["tf32", "f16"])
def test_gemm_hi_lo_exactness(kind):
    """Inputs exactly representable in tf32 / fp16 x small integers: the result must be exact."""
    A = torch.randint(-8, 9, (128, 64)).float()
    B = torch.randint(-8, 9, (64, 32)).float()
    out = _run_gemm([A], [B], [32], 128, kind=kind).cpu()
    assert torch.equal(out, A @ B)


def test_gemm_f16x3_wide_dynamic_range_and_overflow_flag():
    """fp16 split: elements 2^-20 .. 2^3 of the operand scale keep fp32-level accuracy (remainders go subnormal gracefully);
    an operand value beyond the fp16 range raises the status bit instead of passing silently."""
    from jamun_b200 import ops

    gen = torch.Generator().manual_seed(11)
    rows, K, N = 256, 32 * 20, 152
    A = torch.randn(rows, K, generator=gen) * torch.exp2(torch.randint(-20, 4, (rows, K), generator=gen).float())
    B = torch.randn(K, N, generator=gen) * torch.exp2(torch.randint(-12, 1, (K, N), generator=gen).float()) * 1e-2
    out = _run_gemm([A], [B], [160], rows, kind="f16").cpu().double()
    ref = A.double() @ B.double()
    assert (out - ref).abs().max().item() <= 3e-6 * ref.abs().max().item()
    A[5, 7] = 1.0e5
    with pytest.raises(AssertionError):
        _run_gemm([A], [B], [160], rows, kind="f16")


@pytest.mark.parametrize("gemm", ["f16", "tf32"])
@pytest.mark.parametrize("sizes", [[22, 15, 9, 30], [57, 3, 1, 40, 40, 17, 64, 65, 31], [260, 2, 131]])
def test_conv_tc_matches_simt_and_fp64(models, sizes, gemm, monkeypatch):
    """Two-launch tensor-core conv == exact-fp32 SIMT conv (same operands) to ~1e-6, for the initial and a hidden block,
    with the fp16-split (default) and the tf32-split contraction."""
    import kernel_model as KM
    from jamun_b200 import data, engine, ops, synthetic

    monkeypatch.setenv("JAMUN_B200_GEMM", gemm)
    o32, o64, prod = models
    t = synthetic.make_tensors(sizes)
    gen = torch.Generator().manual_seed(1)
    y = (t["pos"] + 0.04 * torch.randn(t["pos"].shape, generator=gen)).cuda()
    topo = engine.Topology(data.Batch.from_tensors(t), "cuda")
    ctx = prod.sigma_context(0.04)
    plan = prod.arch_module.plan(ctx.c_noise, "cuda")
    ybar, p = ops.center_scale(y, topo.chain_ptr, ctx.c_in)
    topo.build_csr(ybar, ctx.r_cut)
    mu, step = plan.radial_grid(ctx.r_cut)
    ops.edge_geom(p, topo.rowptr, topo.col, topo.edst, mu, step, topo.rhat, topo.rb)
    N = p.shape[0]
    for l, x in ((0, None), (1, torch.randn(N, 216, generator=gen).cuda())):
        b = plan.blocks[l]
        if x is None:
            x = torch.randn(N, 56, generator=gen).cuda()
        ops.edge_radial_hidden(topo.rb, topo.ebond, topo.rowptr, b["w0r"], b["b0eff"], topo.h)
        ref = torch.empty(N, 248, device="cuda")
        ops.conv_fwd(x, b["s_in"], b["v_in"], topo.rowptr, topo.col, topo.h, topo.rhat, b["m0"], b["m1"], b["alpha0"], b["alpha1"], ref)
        got = torch.full((N, 248), float("nan"), device="cuda")
        engine.conv_tc(topo, b, x, got)
        vadd = engine.conv_tc_join(topo, b)  # 0e(x)1e->1e part gathered on the side stream, normally added by block_tail
        if vadd is not None:
            got[:, 152:] += vadd
        torch.cuda.synchronize()
        err = (got - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert b["gemm_kind"] == gemm and int(topo.gemm_status.item()) == 0
        assert err <= 5e-5 * max(1.0, scale), f"block {l}: err {err} scale {scale}"  # tcgen05 fp32 accumulation truncates


def test_gemm_column_blocks_and_addend():
    """grid.y column blocks (wide per-node transform) and the epilogue addend."""
    from jamun_b200 import ops, packing

    gen = torch.Generator().manual_seed(3)
    rows, K, N = 300, 128, 480
    rows_pad = 384
    A = torch.randn(rows, K, generator=gen)
    W = torch.randn(K, N, generator=gen)
    add = torch.randn(rows, 32, generator=gen)
    a_op = torch.empty(K // 32 * rows_pad * 32, device="cuda")
    ops.pack_rows(A.cuda(), 0, K, rows_pad, a_op)
    out = torch.full((rows, N), float("nan"), device="cuda")
    img = packing.pack_b_column_blocks(W.cuda(), 160)
    ops.gemm_tf32x3([a_op.data_ptr()], [img.data_ptr()], [K // 32], [160], [160], [0], [1.0], rows, rows_pad, None, out.data_ptr(), N,
                    col_blocks=3, b_block_floats=K // 32 * 2 * 160 * 32)
    ref = A.double() @ W.double()
    assert (out.cpu().double() - ref).abs().max() <= 1e-5 * ref.abs().max()
    out2 = torch.full((rows, 32), float("nan"), device="cuda")
    img2 = packing.pack_b_images(W[:, :32].contiguous().cuda(), 32)
    addc = add.cuda()
    ops.gemm_tf32x3([a_op.data_ptr()], [img2.data_ptr()], [K // 32], [32], [32], [0], [0.5], rows, rows_pad, None, out2.data_ptr(), 32,
                    addend_ptrs=[addc.data_ptr()], addend_ld=[32])
    ref2 = 0.5 * (A.double() @ W[:, :32].double() + add.double())
    assert (out2.cpu().double() - ref2).abs().max() <= 1e-5 * ref2.abs().max()
    # the fp16-split form: column blocks (image stride in 4-byte words is half the tf32 one) and the addend joining the
    # pre-scaled accumulator
    sc = ops.f16_scale(W)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    out3 = torch.full((rows, N), float("nan"), device="cuda")
    img3 = ops.pack_b_f16(W.cuda(), K // 32, 160, sc, n_valid=N, n_inner=N, col_blocks=3)
    ops.gemm_f16x3([a_op.data_ptr()], [img3.data_ptr()], [K // 32], [160], [160], [0], [1.0 / sc], rows, rows_pad, None, out3.data_ptr(), N,
                   col_blocks=3, b_block_floats=K // 32 * 160 * 32, status=status)
    assert (out3.cpu().double() - ref).abs().max() <= 1e-5 * ref.abs().max()
    out4 = torch.full((rows, 32), float("nan"), device="cuda")
    img4 = ops.pack_b_f16(W[:, :32].contiguous().cuda(), K // 32, 32, sc)
    ops.gemm_f16x3([a_op.data_ptr()], [img4.data_ptr()], [K // 32], [32], [32], [0], [0.5 / sc], rows, rows_pad, None, out4.data_ptr(), 32,
                   addend_ptrs=[addc.data_ptr()], addend_ld=[32], addend_scale=[sc], status=status)
    assert (out4.cpu().double() - ref2).abs().max() <= 1e-5 * ref2.abs().max()
    assert int(status.item()) == 0


@pytest.mark.parametrize("sizes", [[22, 15, 9, 30], [57, 3, 1, 40, 40, 17, 64, 65, 31]])
def test_build_tc_matches_ffma_builder(models, sizes, monkeypatch):
    """Tensor-core aggregate builder (3xTF32, tcgen05) writes the same A operand / path-2 sums as the FFMA2 builder."""
    from jamun_b200 import data, engine, ops, synthetic

    o32, o64, prod = models
    t = synthetic.make_tensors(sizes)
    gen = torch.Generator().manual_seed(1)
    y = (t["pos"] + 0.04 * torch.randn(t["pos"].shape, generator=gen)).cuda()
    topo = engine.Topology(data.Batch.from_tensors(t), "cuda")
    ctx = prod.sigma_context(0.04)
    plan = prod.arch_module.plan(ctx.c_noise, "cuda")
    ybar, p = ops.center_scale(y, topo.chain_ptr, ctx.c_in)
    topo.build_csr(ybar, ctx.r_cut)
    mu, step = plan.radial_grid(ctx.r_cut)
    ops.edge_geom(p, topo.rowptr, topo.col, topo.edst, mu, step, topo.rhat, topo.rb)
    N = p.shape[0]
    for l, d_in in ((0, 56), (1, 216)):
        b = plan.blocks[l]
        x = torch.randn(N, d_in, generator=gen).cuda()
        ops.edge_radial_hidden(topo.rb, topo.ebond, topo.rowptr, b["w0r"], b["b0eff"], topo.h)
        res = {}
        if topo.a_ws is None:
            engine.conv_tc(topo, b, x, torch.empty(N, 248, device="cuda"))  # allocates the operand workspace
            engine.conv_tc_join(topo, b)
        for variant in ("ffma", "tc", "tc_tiled"):
            monkeypatch.setenv("JAMUN_B200_BUILD", variant.split("_")[0])
            monkeypatch.setenv("JAMUN_B200_A_LAYOUT", "tile" if variant == "tc_tiled" else "stage")
            out = torch.full((N, 248), float("nan"), device="cuda")
            topo.a_ws.fill_(float("nan"))
            engine.conv_tc(topo, b, x, out)
            vadd = engine.conv_tc_join(topo, b)
            if vadd is not None:
                out[:, 152:] += vadd
            torch.cuda.synchronize()
            nst = 65 * ((d_in > 56) * 11 + (d_in == 56) * 2)
            res[variant] = (topo.a_ws[: nst * topo.chunk_rows * 32].clone(), out)
        a_ref, out_ref = res["ffma"]
        a_tc, out_tc = res["tc"]
        live = ~torch.isnan(a_ref)
        assert torch.equal(torch.isnan(a_tc), ~live), "different set of operand elements written"
        scale = a_ref[live].abs().max().item()
        err = (a_tc[live] - a_ref[live]).abs().max().item()
        assert err <= 2e-6 * max(1.0, scale), f"block {l}: A operand err {err} scale {scale}"
        err_o = (out_tc - out_ref).abs().max().item()
        assert err_o <= 2e-5 * max(1.0, out_ref.abs().max().item()), f"block {l}: conv err {err_o}"
        # the tile-major layout (default of the fp16-split path) holds the same numbers, [row/128][stage][128][32] per segment
        a_tl, out_tl = res["tc_tiled"]
        rp = topo.chunk_rows
        segs = [65 * 2] if d_in == 56 else [65 * 5, 65 * 2, 65 * 2, 65 * 2]
        off = 0
        for nst_seg in segs:
            n = nst_seg * rp * 32
            st_major = a_tc[off:off + n].view(nst_seg, rp // 128, 128, 32).permute(1, 0, 2, 3).reshape(-1)
            tl = a_tl[off:off + n]
            assert torch.equal(torch.isnan(tl), torch.isnan(st_major))
            err_t = (torch.nan_to_num(tl) - torch.nan_to_num(st_major)).abs().max().item()
            assert err_t <= 2e-6 * max(1.0, scale), f"block {l}: tile-major operand err {err_t} scale {scale}"
            off += n
        err_o = (out_tl - out_ref).abs().max().item()
        assert err_o <= 2e-5 * max(1.0, out_ref.abs().max().item()), f"block {l}: conv err {err_o} (tiled)"


def test_block_tail_tc_matches_simt(models, monkeypatch):
    """Block tail as pack -> tcgen05 GEMM -> mix == the exact-fp32 kernel, initial and hidden block, with and without skip/scale."""
    from jamun_b200 import data, engine, ops, synthetic

    o32, o64, prod = models
    t = synthetic.make_tensors([22, 15, 9, 30, 40])
    topo = engine.Topology(data.Batch.from_tensors(t), "cuda")
    ctx = prod.sigma_context(0.04)
    plan = prod.arch_module.plan(ctx.c_noise, "cuda")
    N = topo.N
    gen = torch.Generator().manual_seed(5)
    topo.a_ws = torch.empty(65 * 11 * topo.chunk_rows * 32, device="cuda")
    for l, d_in in ((0, 56), (1, 216), (5, 216)):
        b = plan.blocks[l]
        topo.conv.copy_(torch.randn(N, 248, generator=gen))
        x_in = torch.randn(N, d_in, generator=gen).cuda()
        x_res = torch.randn(N, 216, generator=gen).cuda() if l > 0 else None
        vadd = torch.randn(N, 96, generator=gen).cuda() if l > 0 else None
        skip_w = plan.skips[l - 1] if l > 0 else None
        s_next = plan.scales[l] if l < len(plan.blocks) - 1 else None
        outs = {}
        for impl in ("simt", "tc"):
            monkeypatch.setenv("JAMUN_B200_TAIL", impl)
            x_new = torch.full((N, 216), float("nan"), device="cuda")
            x_sc = torch.full((N, 216), float("nan"), device="cuda") if s_next is not None else None
            engine.block_tail(topo, b, x_in, x_res, skip_w, s_next, x_new, x_sc, vadd)
            torch.cuda.synchronize()
            outs[impl] = (x_new, x_sc)
        for a, c in zip(outs["simt"], outs["tc"]):
            if a is None:
                continue
            assert not torch.isnan(c).any()
            assert (a - c).abs().max().item() <= 2e-6 * max(1.0, a.abs().max().item())


@pytest.mark.parametrize("build", ["tc", "ffma"])
def test_conv_tc_row_chunks(models, monkeypatch, build):
    """Batches larger than the operand workspace are processed in row chunks: chunked == unchunked."""
    from jamun_b200 import data, engine, ops, synthetic

    o32, o64, prod = models
    monkeypatch.setenv("JAMUN_B200_BUILD", build)
    t = synthetic.make_tensors([57, 3, 1, 40, 40, 17, 64, 65, 31, 90])
    gen = torch.Generator().manual_seed(2)
    y = (t["pos"] + 0.04 * torch.randn(t["pos"].shape, generator=gen)).cuda()
    ctx = prod.sigma_context(0.04)
    plan = prod.arch_module.plan(ctx.c_noise, "cuda")
    outs = []
    for chunk in (None, 128):
        topo = engine.Topology(data.Batch.from_tensors(t), "cuda")
        if chunk is not None:
            topo.chunk_rows = chunk
        ybar, p = ops.center_scale(y, topo.chain_ptr, ctx.c_in)
        topo.build_csr(ybar, ctx.r_cut)
        mu, step = plan.radial_grid(ctx.r_cut)
        ops.edge_geom(p, topo.rowptr, topo.col, topo.edst, mu, step, topo.rhat, topo.rb)
        N = p.shape[0]
        assert chunk is None or N > 2 * chunk
        b = plan.blocks[1]
        x = torch.randn(N, 216, generator=torch.Generator().manual_seed(9)).cuda()
        ops.edge_radial_hidden(topo.rb, topo.ebond, topo.rowptr, b["w0r"], b["b0eff"], topo.h)
        got = torch.full((N, 248), float("nan"), device="cuda")
        engine.conv_tc(topo, b, x, got)
        vadd = engine.conv_tc_join(topo, b)
        if vadd is not None:
            got[:, 152:] += vadd
        torch.cuda.synchronize()
        outs.append(got)
    assert not torch.isnan(outs[1]).any()
    assert (outs[0] - outs[1]).abs().max().item() <= 2e-6 * max(1.0, outs[0].abs().max().item())


@pytest.mark.parametrize("k_splits", [2, 5, 16])
def test_gemm_splitk_matches_single_pass(k_splits):
    """Split-K (few row tiles) == the single-pass GEMM up to the accumulation order, and is itself deterministic."""
    from jamun_b200 import ops, packing

    gen = torch.Generator().manual_seed(4)
    rows, rows_pad = 200, 256
    Ks, Ns, pads = [32 * 20, 32 * 7, 32 * 7], [152, 32, 32], [160, 32, 32]
    A = [torch.randn(rows, k, generator=gen) for k in Ks]
    B = [torch.randn(Ks[0], 152, generator=gen), torch.randn(Ks[1], 32, generator=gen)]
    B = [B[0], B[1], B[1]]
    a_ops = []
    for a in A:
        op = torch.empty(a.shape[1] // 32 * rows_pad * 32, device="cuda")
        ops.pack_rows(a.cuda(), 0, a.shape[1], rows_pad, op)
        a_ops.append(op)
    imgs = [packing.pack_b_images(w.cuda(), p) for w, p in zip(B, pads)]
    rs = (torch.rand(rows, generator=gen) + 0.5).cuda()
    args = ([o.data_ptr() for o in a_ops], [i.data_ptr() for i in imgs], [k // 32 for k in Ks], pads, Ns, [0, 152, 184], [0.5, 2.0, 2.0],
            rows, rows_pad, rs.data_ptr())
    ref = torch.full((rows, 216), float("nan"), device="cuda")
    ops.gemm_tf32x3(*args, ref.data_ptr(), 216)
    outs = []
    for _ in range(2):
        out = torch.full((rows, 216), float("nan"), device="cuda")
        partial = torch.full((k_splits * rows * 216,), float("nan"), device="cuda")
        ops.gemm_tf32x3_splitk(*args, out.data_ptr(), 216, k_splits, partial)
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    assert not torch.isnan(outs[0]).any()
    assert (outs[0] - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
