"""Decode what jamun_stage_atb_tc computes on one-hot operands (debug aid for csrc/gemm_atb.cu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from jamun_b200 import ops  # noqa: E402


def run(rows, n_stages, W, A, B, impl):
    rows_pad = (rows + 127) // 128 * 128
    a_op = torch.zeros(1, n_stages, rows_pad, 32)
    a_op[:, :, :rows] = A.reshape(1, n_stages, 32, rows).permute(0, 1, 3, 2)
    rr = torch.arange(rows_pad)[:, None]
    ll = torch.arange(32)[None, :]
    pos = (((ll // 4) ^ (rr % 8)) * 4 + ll % 4).expand(1, n_stages, rows_pad, 32)
    a_sw = torch.empty_like(a_op).scatter_(3, pos, a_op).cuda().contiguous()
    out = torch.zeros(n_stages, W, 32, device="cuda")
    ops.ATB_IMPL = impl
    ops.stage_atb_auto(a_sw.data_ptr(), n_stages * rows_pad * 32, 1, n_stages, 1, rows, rows_pad, B.cuda().contiguous(), 0, W, W, out, 1, W)
    torch.cuda.synchronize()
    return out.cpu().permute(0, 2, 1).reshape(n_stages * 32, W)  # [m, w]


torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
for rows, n_stages, W in [(32, 4, 32), (64, 4, 32), (64, 8, 64), (300, 8, 152)]:
    for rstar in (0, 1, 9, rows - 1):
        A = torch.zeros(n_stages * 32, rows)
        A[:, rstar] = torch.arange(n_stages * 32) + 1.0
        B = torch.zeros(rows, W)
        B[rstar] = torch.arange(W) + 1.0
        want = A @ B
        for impl in ("simt", "tc"):
            got = run(rows, n_stages, W, A, B, impl)
            err = (got - want).abs().max().item()
            print(f"rows {rows} stages {n_stages} W {W} r* {rstar} {impl}: max err {err:.3g}")
            if err > 1e-3 and impl == "tc":
                nz = got.nonzero()
                print("  nonzero count", len(nz), "of expected", int((want != 0).sum()))
                print("  got[0:4, 0:10]\n", got[0:4, 0:10])
                print("  got[32:36, 0:10]\n", got[32:36, 0:10])
                if len(nz):
                    print("  first nonzeros", nz[:8].tolist(), [round(got[i, j].item(), 1) for i, j in nz[:8].tolist()])
    g = torch.Generator().manual_seed(1)
    A = torch.randn(n_stages * 32, rows, generator=g)
    B = torch.randn(rows, W, generator=g)
    want = A.double() @ B.double()
    for impl in ("simt", "tc"):
        got = run(rows, n_stages, W, A, B, impl)
        print(f"rows {rows} stages {n_stages} W {W} random {impl}: max err {(got - want).abs().max().item():.3g} of {want.abs().max().item():.3g}")
