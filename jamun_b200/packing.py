"""Host-side re-layout of weight matrices into the byte images the tensor-core GEMM streams with 1-D bulk copies.

A stage of the B operand is a [n_pad x 32] tile (n_pad output columns, 32 consecutive K rows) stored K-major in the
UMMA SWIZZLE_128B shared-memory layout: row n occupies 128 bytes at (n // 8) * 1024 + (n % 8) * 128 and its eight
16-byte chunks are XOR-permuted with (n % 8).  Each stage image is the tf32 "hi" tile followed by the exact fp32
remainder "lo" tile (3xTF32 split: w = hi + lo, hi = w & 0xFFFFE000).
"""
from __future__ import annotations

import torch


def _swizzle_positions(n_pad: int, device) -> torch.Tensor:
    n = torch.arange(n_pad, device=device)[:, None]
    kk = torch.arange(32, device=device)[None, :]
    byte = (n // 8) * 1024 + (n % 8) * 128 + (((kk // 4) ^ (n % 8)) * 16) + (kk % 4) * 4
    return (byte // 4).reshape(-1)  # float index inside one [n_pad x 32] image


def split_tf32(w: torch.Tensor):
    hi = (w.contiguous().view(torch.int32) & -8192).view(torch.float32)  # 0xFFFFE000
    return hi, w - hi


def pack_b_images(w: torch.Tensor, n_pad: int) -> torch.Tensor:
    """w: [K, N] fp32 with K % 32 == 0 and N <= n_pad  ->  [K/32, 2, n_pad*32] swizzled (hi, lo) stage images."""
    K, N = w.shape
    assert K % 32 == 0 and N <= n_pad and n_pad % 16 == 0
    S = K // 32
    wp = torch.zeros(K, n_pad, dtype=torch.float32, device=w.device)
    wp[:, :N] = w
    tiles = wp.reshape(S, 32, n_pad).permute(0, 2, 1).reshape(S, n_pad * 32)  # [stage][(n, kk)]
    pos = _swizzle_positions(n_pad, w.device)
    hi, lo = split_tf32(tiles)
    out = torch.zeros(S, 2, n_pad * 32, dtype=torch.float32, device=w.device)
    out[:, 0, pos] = hi
    out[:, 1, pos] = lo
    return out.contiguous()


def conv_k_layout(m0: torch.Tensor, m1: torch.Tensor, s_in: int, v_in: int):
    """Re-order the conv contraction weights m0 [65,U0,152], m1 [65,U1,32] into the K order of jamun_conv_build_a
    (per radial channel: scalar slots zero-padded to 32, then the vector slots) -> (w0 [K0,152], w1 [K1,32])."""
    ns = (s_in + 31) // 32
    K = m0.shape[0]

    def pad_scalar(m):
        out = torch.zeros(K, ns * 32, m.shape[2], dtype=m.dtype, device=m.device)
        out[:, :s_in] = m[:, :s_in]
        return out

    w0 = [pad_scalar(m0)]
    w1 = [pad_scalar(m1)]
    if v_in:
        assert v_in == 32
        w0.append(m0[:, s_in:s_in + v_in])
        w1 += [m1[:, s_in:s_in + v_in], m1[:, s_in + v_in:s_in + 2 * v_in]]
    w0 = torch.cat(w0, dim=1)
    w1 = torch.cat(w1, dim=1)
    return w0.reshape(-1, m0.shape[2]).contiguous(), w1.reshape(-1, m1.shape[2]).contiguous()
