"""Debug aid: accuracy of jamun_gemm_tf32x3 against fp64 over a few shapes (growth of the tensor-core accumulation error with K)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_tc import _run_gemm
for rows, K, N, n_pad in [(1000, 1280, 16, 16), (1000, 1280, 32, 32), (256, 1280, 16, 16), (128, 32, 16, 16), (128, 32, 48, 48), (512, 9984, 152, 160), (512, 12480, 32, 32)]:
    gen = torch.Generator().manual_seed(1)
    A = torch.randn(rows, K, generator=gen); B = torch.randn(K, N, generator=gen) * 0.3
    out = _run_gemm([A], [B], [n_pad], rows).cpu().double()
    ref = A.double() @ B.double()
    ref32 = (A.cuda() @ B.cuda()).cpu().double()
    mag = (A.abs().double() @ B.abs().double())
    err = (out - ref).abs()
    print(f"rows={rows} K={K} N={N}: maxerr={err.max():.3e} scale={ref.abs().max():.3e} err/sum|ab|={(err/mag).max():.3e} mean_signed={(out-ref).mean():.3e} torch_fp32_err={(ref32-ref).abs().max():.3e} nan={int(torch.isnan(out).sum())} badrows={(err.max(1).values > 1e-3).nonzero().flatten()[:10].tolist()}")
