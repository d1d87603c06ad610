"""Debug aid: compare the FFMA2 and tensor-core aggregate builders element by element."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(__file__))
from jamun_b200 import data, engine, ops, synthetic, factory

sizes = [22, 15, 9, 30]
prod = factory.default_denoiser().cuda()
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
rp = topo.chunk_rows
print("N", N, "rows_pad", rp, "max_degree", topo.max_degree)
for l, d_in in ((0, 56), (1, 216)):
    b = plan.blocks[l]
    x = torch.randn(N, d_in, generator=gen).cuda()
    ops.edge_radial_hidden(topo.rb, topo.ebond, topo.rowptr, b["w0r"], b["b0eff"], topo.h)
    res = {}
    for variant in ("1", "20"):
        os.environ["JAMUN_BUILD_VARIANT"] = variant
        out = torch.full((N, 248), float("nan"), device="cuda")
        if topo.a_ws is not None:
            topo.a_ws.fill_(float("nan"))
        engine.conv_tc(topo, b, x, out)
        torch.cuda.synchronize()
        nst = 65 * (11 if d_in > 56 else 2)
        res[variant] = (topo.a_ws[: nst * rp * 32].clone().view(nst, rp, 32), out)
    a_ref, o_ref = res["1"]
    a_tc, o_tc = res["20"]
    nr, nt = torch.isnan(a_ref), torch.isnan(a_tc)
    print(f"block {l}: stages {a_ref.shape[0]}; nan ref {nr.sum().item()} tc {nt.sum().item()} mismatch {(nr != nt).sum().item()}")
    mm = (nr != nt)
    if mm.any():
        st = mm.any(dim=2).any(dim=1).nonzero().flatten()
        print("  stages with mismatch:", st[:40].tolist(), "...", len(st))
        rows = mm.any(dim=2).any(dim=0).nonzero().flatten()
        print("  rows with mismatch:", rows[:40].tolist(), "...", len(rows))
        pos = mm.any(dim=1).any(dim=0).nonzero().flatten()
        print("  positions with mismatch:", pos.tolist())
    both = ~nr & ~nt
    d = (a_ref - a_tc).abs()
    d[~both] = 0
    print("  max abs err", d.max().item(), "scale", a_ref[~nr].abs().max().item())
    if d.max() > 1e-4:
        st = (d.amax(dim=(1, 2)) > 1e-4).nonzero().flatten()
        print("  bad stages", st[:60].tolist(), len(st))
        rows = (d.amax(dim=(0, 2)) > 1e-4).nonzero().flatten()
        print("  bad rows", rows[:60].tolist(), len(rows))
        s0 = st[0].item(); r0 = rows[0].item()
        print("  ref", a_ref[s0, r0].tolist()); print("  tc ", a_tc[s0, r0].tolist())
    print("  out err", (o_ref - o_tc).abs().max().item(), "scale", o_ref.abs().max().item())
