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
    """Re-order the conv contraction weights m0 [65,U0,152], m1 [65,U1,32] into the operands of the tensor-core path:
    w0 [K0,152]  : 0e contraction in the K order of jamun_conv_build_a (per radial channel: scalar slots zero-padded to
                   32, then the x_v.rhat slot);
    w1 [K1,32]   : 1e contraction of the aggregated vector paths (per radial channel: x_v[c] slot, cross slot); None if v_in == 0;
    wy [32*NS, 65*32] : per-node transform of the path 0e(x)1e->1e, wy[u, k'*32+w] = m1[k', u, w]."""
    ns = (s_in + 31) // 32
    K = m0.shape[0]
    pad = torch.zeros(K, ns * 32, m0.shape[2], dtype=m0.dtype, device=m0.device)
    pad[:, :s_in] = m0[:, :s_in]
    w0 = [pad]
    w1 = None
    if v_in:
        assert v_in == 32
        w0.append(m0[:, s_in:s_in + v_in])
        w1 = torch.cat([m1[:, s_in:s_in + v_in], m1[:, s_in + v_in:s_in + 2 * v_in]], dim=1).reshape(-1, m1.shape[2]).contiguous()
    w0 = torch.cat(w0, dim=1).reshape(-1, m0.shape[2]).contiguous()
    wy = torch.zeros(ns * 32, K * m1.shape[2], dtype=m1.dtype, device=m1.device)
    wy[:s_in] = m1[:, :s_in].permute(1, 0, 2).reshape(s_in, -1)
    return w0, w1, wy


def pack_b_column_blocks(w: torch.Tensor, block: int = 160) -> torch.Tensor:
    """w [K, N] with N % block == 0 -> [N/block, K/32, 2, block*32]: one pack_b_images per column block."""
    K, N = w.shape
    assert N % block == 0
    return torch.stack([pack_b_images(w[:, c:c + block].contiguous(), block) for c in range(0, N, block)]).contiguous()


# ---- device-side packing (jamun_pack_b): the same images, produced by one kernel launch per operand ---------------------
_ROW_MAPS = {}


def conv_row_maps(s_in: int, v_in: int, device):
    """int32 row maps for jamun_pack_b reproducing conv_k_layout without materialising w0 / w1:
    map0[k] -> row of m0.view(65*(s_in+v_in), 152) (or -1 for the zero padding of the scalar slots),
    map1[k] -> row of m1.view(65*(s_in+2 v_in), 32) (None if v_in == 0)."""
    key = (s_in, v_in, str(device))
    hit = _ROW_MAPS.get(key)
    if hit is None:
        ns = (s_in + 31) // 32
        u0, u1 = s_in + v_in, s_in + 2 * v_in
        per = ns * 32 + v_in
        kp = torch.arange(65)[:, None]
        r = torch.arange(per)[None, :]
        m = torch.where(r < s_in, kp * u0 + r, torch.where(r < ns * 32, torch.full_like(r, -1), kp * u0 + s_in + (r - ns * 32)))
        map0 = m.reshape(-1).to(torch.int32)
        map1 = None
        if v_in:
            r1 = torch.arange(2 * v_in)[None, :]
            map1 = (kp * u1 + s_in + r1).reshape(-1).to(torch.int32).to(device)
        hit = _ROW_MAPS[key] = (map0.to(device), map1)
    return hit


def pack_conv_operands_device(m0: torch.Tensor, m1: torch.Tensor, s_in: int, v_in: int, f16_scales=None):
    """(b0_img, b1_img | None, wy_img) for CUDA tensors m0 [65, s_in+v_in, 152], m1 [65, s_in+2 v_in, 32] -- three launches
    of jamun_pack_b.  Equal, bit for bit, to pack_b_images / pack_b_column_blocks of conv_k_layout(m0, m1, s_in, v_in).
    f16_scales = (scale of m0, scale of m1): the fp16-split images of jamun_gemm_f16x3 instead (jamun_pack_b_f16)."""
    from . import ops

    ns = (s_in + 31) // 32
    map0, map1 = conv_row_maps(s_in, v_in, m0.device)
    K = m0.shape[0]
    u0, u1 = s_in + v_in, s_in + 2 * v_in
    if f16_scales is not None:
        sc0, sc1 = f16_scales
        b0 = ops.pack_b_f16(m0.reshape(K * u0, m0.shape[2]), K * (ns * 32 + v_in) // 32, 160, sc0, row_map=map0)
        b1 = ops.pack_b_f16(m1.reshape(K * u1, m1.shape[2]), K * 2 * v_in // 32, 32, sc1, row_map=map1) if v_in else None
        wy = ops.pack_b_f16(m1.reshape(K * u1, m1.shape[2]), ns, 128, sc1, k_src=s_in, n_valid=K * m1.shape[2], n_inner=m1.shape[2],
                            outer_rows=u1, col_blocks=17)
        return b0, b1, wy
    b0 = ops.pack_b(m0.reshape(K * u0, m0.shape[2]), n_stages=K * (ns * 32 + v_in) // 32, n_pad=160, row_map=map0)
    b1 = ops.pack_b(m1.reshape(K * u1, m1.shape[2]), n_stages=K * 2 * v_in // 32, n_pad=32, row_map=map1) if v_in else None
    # W_y[u, k'*32 + w] = m1[k', u, w] for the scalar rows u < s_in; 65*32 = 2080 columns in 17 blocks of 128
    wy = ops.pack_b(m1.reshape(K * u1, m1.shape[2]), n_stages=ns, n_pad=128, k_src=s_in, n_valid=K * m1.shape[2],
                    n_inner=m1.shape[2], outer_rows=u1, col_blocks=17)
    return b0, b1, wy
