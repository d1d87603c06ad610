"""Debug: dump per-item role timestamps of the tensor-core builder (CTA 0)."""
import ctypes, os, sys
os.environ["JAMUN_B200_BUILD"] = "tc"; os.environ["JAMUN_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jamun_b200 import _lib, data, engine, factory, ops, synthetic
prod = factory.default_denoiser().cuda()
WL = sys.argv[1] if len(sys.argv) > 1 else "2AA"
t = synthetic.make_tensors(synthetic.workload_sizes(WL, 64 if WL == "protein1000" else 1024), n_res={"2AA": 2, "4AA": 4, "protein1000": 100}[WL])
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
b = plan.blocks[1]
x = torch.randn(N, 216, generator=gen).cuda()
ops.edge_radial_hidden(topo.rb, topo.ebond, topo.rowptr, b["w0r"], b["b0eff"], topo.h)
out = torch.empty(N, 248, device="cuda")
for _ in range(3):
    engine.conv_tc(topo, b, x, out)
    engine.conv_tc_join(topo, b)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (64 * 16))()
lib = _lib.lib()
lib.jamun_debug_tc_trace.argtypes = [ctypes.c_void_p]
assert lib.jamun_debug_tc_trace(buf) == 0
rows = [[buf[i * 16 + k] for k in range(16)] for i in range(64)]
t0 = rows[0][0]
names = ["P.top", "P.ld", "P.emp", "P.arr", "M.top", "M.full", "M.temp", "M.done", "E.top", "E.f0", "E.f1", "E.f2", "E.d0", "E.d1", "E.d2"]
print("item " + " ".join(f"{n:>7s}" for n in names))
for i in range(20, 36):
    print(f"{i:4d} " + " ".join(f"{rows[i][k] - t0:7d}" if rows[i][k] else "      -" for k in range(15)))
