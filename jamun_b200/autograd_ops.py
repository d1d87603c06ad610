"""torch.library operator layer of the training path ("Python wrappers registered with torch.library.custom_op
+ register_fake + register_autograd").

Every operator here is a sequence of this library's CUDA kernels behind the C ABI (jamun_b200.ops): the forward ops are the
sampling path's kernels (tcgen05 conv, fused block tail, head), the ``*_bwd`` ops are the hand-written backward kernels of
csrc/train_*.cu.  torch contributes the operator registry, fake-tensor shapes, the autograd tape and device memory -- no
arithmetic.  Reference semantics: /root/reference/src/jamun/e3tools/nn/_conv.py:93-119 (Conv), _conv.py:204-221 +
_interaction.py:26-30 + _gate.py:63 + model/arch/e3conv.py:132-133 (block tail), _mlp.py:37-114 (head), _mlp.py:10-34 (radial
MLP), model/atom_embedding.py:58-76, model/noise_conditioning.py:27-73, model/denoiser.py:200-287 (loss), utils/align.py:9-56.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import ops, packing

S, V, HID, GIN, SO = ops.S, ops.V, ops.HID, ops.GATE_IN, ops.S + ops.V
Y_LD = 17 * 128
NS_LIB = "jamun_b200"


def _rows_pad(n: int) -> int:
    return (n + 127) // 128 * 128


def _empty(*shape, like: Tensor, dtype=torch.float32) -> Tensor:
    return torch.empty(*shape, dtype=dtype, device=like.device)


class _Scratch:  # the split-K partial buffer engine._contract caches on its owner
    def __init__(self, device):
        self.device = device
        self.gemm_partial = None


_scratch = {}


def _scratch_of(device) -> _Scratch:
    if device not in _scratch:
        _scratch[device] = _Scratch(device)
    return _scratch[device]


def _shape(s_in: int, v_in: int):
    ns = (s_in + 31) // 32
    nsl0, nsl1 = ns + (1 if v_in else 0), (2 if v_in else 0)
    return ns, nsl0, nsl1


def _node_transform(x: Tensor, wy_img: Tensor, s_in: int, rows_pad: int) -> Tensor:
    """Y = x_s . W_y [N, 2176] (per-node transform of the path 0e(x)1e->1e) on the tensor cores."""
    N = x.shape[0]
    ns = (s_in + 31) // 32
    xs_op = _empty(ns * rows_pad * 32, like=x)
    ops.pack_rows(x, 0, s_in, rows_pad, xs_op)
    y = _empty(N, Y_LD, like=x)
    ops.gemm_tf32x3([xs_op.data_ptr()], [wy_img.data_ptr()], [ns], [128], [128], [0], [1.0], N, rows_pad, None, y.data_ptr(), Y_LD,
                    col_blocks=17, b_block_floats=ns * 2 * 128 * 32)
    return y


def _build_aggregate(x, s_in, v_in, rowptr, col, h, rhat, rows_pad):
    """The contraction operand A (jamun_conv_build_tc) and 1/deg; returns (a_ws, a1_offset, comp_stride, inv_deg)."""
    N = x.shape[0]
    _, nsl0, nsl1 = _shape(s_in, v_in)
    st0, st1 = 65 * nsl0, 65 * nsl1
    a_ws = _empty((st0 + 3 * st1) * rows_pad * 32, like=x)
    a1_off, comp = st0 * rows_pad * 32, st1 * rows_pad * 32
    inv_deg = _empty(N, like=x)
    base = a_ws.data_ptr()
    ops.conv_build_tc(x, s_in, v_in, rowptr, col, h, rhat, 0, N, rows_pad, base, base + 4 * a1_off if v_in else None, comp, inv_deg)
    return a_ws, a1_off, comp, inv_deg


# =====================================================================================================================
# conv
# =====================================================================================================================
def _conv_forward(x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, s_in, v_in, alpha0, alpha1):
    """-> (out, a_ws, inv_deg, y): the conv output and the operands its backward can reuse."""
    from . import engine

    x, h, m0, m1 = x.contiguous(), h.contiguous(), m0.contiguous(), m1.contiguous()
    N = x.shape[0]
    out = _empty(N, GIN, like=x)
    if N == 0:
        return out, None, None, None
    rows_pad = _rows_pad(N)
    _, nsl0, nsl1 = _shape(s_in, v_in)
    st0, st1 = 65 * nsl0, 65 * nsl1
    b0, b1, wy = packing.pack_conv_operands_device(m0, m1, s_in, v_in)
    y = _node_transform(x, wy, s_in, rows_pad)
    a_ws, a1_off, comp, inv_deg = _build_aggregate(x, s_in, v_in, rowptr, col, h, rhat, rows_pad)
    base = a_ws.data_ptr()
    t_edge = _empty(col.numel(), 32, like=x)
    sc = _scratch_of(x.device)
    if v_in:
        a_ptrs = [base] + [base + 4 * (a1_off + c * comp) for c in range(3)]
        engine._contract(sc, a_ptrs, [b0.data_ptr()] + [b1.data_ptr()] * 3, [st0, st1, st1, st1], [160, 32, 32, 32],
                         [152, 32, 32, 32], [0, 152, 184, 216], [alpha0, alpha1, alpha1, alpha1], N, rows_pad, inv_deg.data_ptr(),
                         out.data_ptr())
        p2 = _empty(N, 96, like=x)
        ops.conv_p2(rowptr, src_rowptr, src_eid, h, rhat, y, t_edge, p2.data_ptr(), 96, alpha1)
        ops.add_cols(out, SO, p2, 96)
    else:
        engine._contract(sc, [base], [b0.data_ptr()], [st0], [160], [152], [0], [alpha0], N, rows_pad, inv_deg.data_ptr(),
                         out.data_ptr())
        ops.conv_p2(rowptr, src_rowptr, src_eid, h, rhat, y, t_edge, out.data_ptr() + 4 * SO, GIN, alpha1)
    return out, a_ws, inv_deg, y




@torch.library.custom_op(f"{NS_LIB}::conv", mutates_args=())
def conv(x: Tensor, h: Tensor, rhat: Tensor, rowptr: Tensor, col: Tensor, edst: Tensor, src_rowptr: Tensor, src_eid: Tensor,
         m0: Tensor, m1: Tensor, s_in: int, v_in: int, alpha0: float, alpha1: float) -> Tensor:
    """Conv.forward: x [N, s_in+3 v_in] (SoA) -> [N, 248]; h [cap, 64] radial hidden, CSR by receiver and by source."""
    return _conv_forward(x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, s_in, v_in, alpha0, alpha1)[0]


@conv.register_fake
def _(x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, s_in, v_in, alpha0, alpha1):
    return x.new_empty(x.shape[0], GIN)


@torch.library.custom_op(f"{NS_LIB}::conv_bwd", mutates_args=())
def conv_bwd(dout: Tensor, x: Tensor, h: Tensor, rhat: Tensor, rowptr: Tensor, col: Tensor, edst: Tensor, src_rowptr: Tensor,
             src_eid: Tensor, m0: Tensor, m1: Tensor, s_in: int, v_in: int, alpha0: float, alpha1: float,
             a_saved: Optional[Tensor] = None, inv_deg_saved: Optional[Tensor] = None, y_saved: Optional[Tensor] = None
             ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (dx [N, D_in], dh [cap, 64], dm0, dm1); scheme in csrc/train_conv_bwd.cu.  a_saved / inv_deg_saved / y_saved: the
    forward's aggregate operand, 1/deg and per-node transform when conv_keep kept them (otherwise recomputed here)."""
    dout, x, h, m0, m1 = dout.contiguous(), x.contiguous(), h.contiguous(), m0.contiguous(), m1.contiguous()
    N, D_in, cap = x.shape[0], x.shape[1], col.numel()
    ns, nsl0, nsl1 = _shape(s_in, v_in)
    u0, u1 = s_in + v_in, s_in + 2 * v_in
    dx = _empty(N, D_in, like=x)
    dh = torch.zeros(cap, 64, dtype=torch.float32, device=x.device)
    dm0, dm1 = torch.empty_like(m0), torch.empty_like(m1)
    if N == 0:
        return dx, dh, dm0.zero_(), dm1.zero_()
    rows_pad = _rows_pad(N)
    # the forward operands Y (per-node transform) and A (aggregate): kept by conv_keep, or recomputed
    if y_saved is not None and y_saved.numel():
        y = y_saved
    else:
        wy = ops.pack_b(m1.reshape(65 * u1, 32), n_stages=ns, n_pad=128, k_src=s_in, n_valid=65 * 32, n_inner=32, outer_rows=u1,
                        col_blocks=17)
        y = _node_transform(x, wy, s_in, rows_pad)
    if a_saved is not None and a_saved.numel():
        a_ws, inv_deg = a_saved, inv_deg_saved
        a1_off, comp = 65 * nsl0 * rows_pad * 32, 65 * nsl1 * rows_pad * 32
    else:
        a_ws, a1_off, comp, inv_deg = _build_aggregate(x, s_in, v_in, rowptr, col, h, rhat, rows_pad)
    base = a_ws.data_ptr()
    g = _empty(N, GIN, like=x)
    ops.conv_bwd_scale(dout, inv_deg, alpha0, alpha1, g)
    # dM = A^T . G over the nodes
    rows0 = [32] * ns
    rows0[-1] = s_in - 32 * (ns - 1)
    slot_row0 = [32 * s for s in range(ns)] + ([s_in] if v_in else [])
    slot_rows = rows0 + ([v_in] if v_in else [])
    ops.stage_atb_auto(base, 0, 1, 65 * nsl0, nsl0, N, rows_pad, g, 0, 0, SO, dm0, 0, u0, slot_row0, slot_rows)
    if v_in:
        ops.stage_atb_auto(base + 4 * a1_off, comp, 3, 65 * 2, 2, N, rows_pad, g, SO, V, V, dm1, 0, u1, [s_in, s_in + v_in], [v_in, v_in])
    # dA = G . M^T (column blocks of 128 over the (k', u') index, same order as the forward K axis)
    map0, map1 = packing.conv_row_maps(s_in, v_in, x.device)
    k0 = 65 * nsl0 * 32
    cb0 = (k0 + 127) // 128
    g_op = _empty(5 * rows_pad * 32, like=x)
    ops.pack_rows(g, 0, SO, rows_pad, g_op)
    bt0 = ops.pack_b(m0.reshape(65 * u0, SO), n_stages=5, n_pad=128, row_map=map0, k_src=k0, n_valid=SO, col_blocks=cb0,
                     transpose=True)
    dA0 = _empty(N, cb0 * 128, like=x)
    ops.gemm_tf32x3([g_op.data_ptr()], [bt0.data_ptr()], [5], [128], [128], [0], [1.0], N, rows_pad, None, dA0.data_ptr(),
                    cb0 * 128, col_blocks=cb0, b_block_floats=5 * 2 * 128 * 32)
    dA1 = None
    if v_in:
        k1 = 65 * 64
        cb1 = (k1 + 127) // 128
        bt1 = ops.pack_b(m1.reshape(65 * u1, V), n_stages=1, n_pad=128, row_map=map1, k_src=k1, n_valid=V, col_blocks=cb1,
                         transpose=True)
        dA1 = _empty(3, N, cb1 * 128, like=x)
        g1_op = _empty(rows_pad * 32, like=x)
        for c in range(3):
            ops.pack_rows(g, SO + V * c, V, rows_pad, g1_op)
            ops.gemm_tf32x3([g1_op.data_ptr()], [bt1.data_ptr()], [1], [128], [128], [0], [1.0], N, rows_pad, None,
                            dA1[c].data_ptr(), cb1 * 128, col_blocks=cb1, b_block_floats=2 * 128 * 32)
    # per-edge gradients of the aggregated paths
    dxe = _empty(cap, D_in, like=x)
    ops.conv_bwd_edge(x, s_in, v_in, rowptr, col, h, rhat, dA0, dA1, dh, dxe)
    # path 0e(x)1e->1e: dY (source-major), dM2 = x_s^T . dY, dx_s += dY . M2^T
    dy_op = _empty(65 * rows_pad * 32, like=x)
    ops.conv_bwd_p2(src_rowptr, src_eid, edst, h, rhat, y, g, rows_pad, dh, dy_op)
    ops.stage_atb_auto(dy_op.data_ptr(), 0, 1, 65, 1, N, rows_pad, x, 0, 0, s_in, dm1, 1, u1)
    m2t = m1[:, :s_in, :].permute(1, 0, 2).reshape(s_in, 65 * V).contiguous()  # [u, (k', w)] (re-layout only)
    bt2 = ops.pack_b(m2t, n_stages=65, n_pad=128, k_src=s_in, n_valid=65 * V, transpose=True)
    dxs2 = _empty(N, 128, like=x)
    ops.gemm_tf32x3([dy_op.data_ptr()], [bt2.data_ptr()], [65], [128], [128], [0], [1.0], N, rows_pad, None, dxs2.data_ptr(), 128)
    ops.conv_bwd_gather(src_rowptr, src_eid, dxe, dxs2, s_in, dx)
    return dx, dh, dm0, dm1


@conv_bwd.register_fake
def _(dout, x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, s_in, v_in, alpha0, alpha1, a_saved=None, inv_deg_saved=None,
      y_saved=None):
    return torch.empty_like(x), h.new_empty(col.numel(), 64), torch.empty_like(m0), torch.empty_like(m1)


def _conv_setup(ctx, inputs, output):
    x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, s_in, v_in, alpha0, alpha1 = inputs
    ctx.save_for_backward(x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1)
    ctx.consts = (s_in, v_in, alpha0, alpha1)


def _conv_backward(ctx, dout):
    x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1 = ctx.saved_tensors
    dx, dh, dm0, dm1 = torch.ops.jamun_b200.conv_bwd(dout, x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, *ctx.consts)
    return dx, dh, None, None, None, None, None, None, dm0, dm1, None, None, None, None


conv.register_autograd(_conv_backward, setup_context=_conv_setup)


# ---- conv_keep: the same forward, keeping the aggregate operand A (3.2 GB per hidden layer at 35 k atoms), 1/deg and the per-node
# transform Y for the backward instead of rebuilding them there (builder + transform recompute: ~5 ms of a 79 ms step).  A layer
# whose operand exceeds JAMUN_B200_TRAIN_KEEP_A_GB (default 4 GB, i.e. <= 24 GB over the six layers) is recomputed as before.
def _keep_limit_bytes() -> float:
    return float(os.environ.get("JAMUN_B200_TRAIN_KEEP_A_GB", "4")) * 2 ** 30


def _a_floats(N: int, s_in: int, v_in: int) -> int:
    _, nsl0, nsl1 = _shape(s_in, v_in)
    return (65 * nsl0 + 3 * 65 * nsl1) * _rows_pad(N) * 32


@torch.library.custom_op(f"{NS_LIB}::conv_keep", mutates_args=())
def conv_keep(x: Tensor, h: Tensor, rhat: Tensor, rowptr: Tensor, col: Tensor, edst: Tensor, src_rowptr: Tensor, src_eid: Tensor,
              m0: Tensor, m1: Tensor, s_in: int, v_in: int, alpha0: float, alpha1: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (out [N, 248], a_ws, inv_deg, y); the last three are empty when the operand is too large to keep."""
    out, a_ws, inv_deg, y = _conv_forward(x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, s_in, v_in, alpha0, alpha1)
    if a_ws is None or 4 * a_ws.numel() > _keep_limit_bytes():
        return out, _empty(0, like=out), _empty(0, like=out), _empty(0, like=out)
    return out, a_ws, inv_deg, y


@conv_keep.register_fake
def _(x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, s_in, v_in, alpha0, alpha1):
    N = x.shape[0]
    if N == 0 or 4 * _a_floats(N, s_in, v_in) > _keep_limit_bytes():
        return x.new_empty(N, GIN), x.new_empty(0), x.new_empty(0), x.new_empty(0)
    return x.new_empty(N, GIN), x.new_empty(_a_floats(N, s_in, v_in)), x.new_empty(N), x.new_empty(N, Y_LD)


def _conv_keep_setup(ctx, inputs, output):
    x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, s_in, v_in, alpha0, alpha1 = inputs
    _, a_ws, inv_deg, y = output
    ctx.save_for_backward(x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, a_ws, inv_deg, y)
    ctx.consts = (s_in, v_in, alpha0, alpha1)
    ctx.set_materialize_grads(False)  # no zero-filled 3.2 GB "gradients" for the kept operands, which nothing differentiates


def _conv_keep_backward(ctx, dout, da, dinv, dy):
    x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, a_ws, inv_deg, y = ctx.saved_tensors
    if dout is None:
        dout = torch.zeros(x.shape[0], GIN, dtype=x.dtype, device=x.device)
    dx, dh, dm0, dm1 = torch.ops.jamun_b200.conv_bwd(dout, x, h, rhat, rowptr, col, edst, src_rowptr, src_eid, m0, m1, *ctx.consts,
                                                     a_ws, inv_deg, y)
    return dx, dh, None, None, None, None, None, None, dm0, dm1, None, None, None, None


conv_keep.register_autograd(_conv_keep_backward, setup_context=_conv_keep_setup)


# =====================================================================================================================
# block tail: Gate -> self-interaction + skip Linear -> noise-conditional skip / next-layer scaling
# =====================================================================================================================
@torch.library.custom_op(f"{NS_LIB}::block_tail", mutates_args=())
def block_tail(conv_out: Tensor, x_in: Tensor, x_res: Optional[Tensor], wself_s: Tensor, wself_v: Tensor, wskip_s: Tensor,
               wskip_v: Optional[Tensor], skip_w: Optional[Tensor], s_next: Optional[Tensor], s_in: int, v_in: int, c_act: float,
               c_gate: float) -> Tuple[Tensor, Tensor]:
    """-> (x_new, x_scaled); x_scaled is empty when s_next is None (last block)."""
    N = conv_out.shape[0]
    x_new = _empty(N, HID, like=conv_out)
    x_scaled = _empty(N, HID, like=conv_out) if s_next is not None else _empty(0, like=conv_out)
    if N:
        ops.block_tail(conv_out.contiguous(), x_in.contiguous(), s_in, v_in, None if x_res is None else x_res.contiguous(),
                       wself_s.contiguous(), wself_v.contiguous(), wskip_s.contiguous(),
                       None if wskip_v is None else wskip_v.contiguous(), skip_w, s_next, c_act, c_gate, x_new,
                       x_scaled if s_next is not None else None)
    return x_new, x_scaled


@block_tail.register_fake
def _(conv_out, x_in, x_res, wself_s, wself_v, wskip_s, wskip_v, skip_w, s_next, s_in, v_in, c_act, c_gate):
    N = conv_out.shape[0]
    return conv_out.new_empty(N, HID), conv_out.new_empty((N, HID) if s_next is not None else (0,))


@torch.library.custom_op(f"{NS_LIB}::block_tail_bwd", mutates_args=())
def block_tail_bwd(dx_new: Optional[Tensor], dx_scaled: Optional[Tensor], conv_out: Tensor, x_in: Tensor, x_res: Optional[Tensor],
                   wself_s: Tensor, wself_v: Tensor, wskip_s: Tensor, wskip_v: Optional[Tensor], skip_w: Optional[Tensor],
                   s_next: Optional[Tensor], s_in: int, v_in: int, c_act: float, c_gate: float) -> List[Tensor]:
    """-> [dconv, dx_in, dx_res, dwself_s, dwself_v, dwskip_s, dwskip_v, dskip_w, ds_next] (empty tensors where not applicable)."""
    conv_out, x_in = conv_out.contiguous(), x_in.contiguous()
    N, D_in = conv_out.shape[0], x_in.shape[1]
    like = conv_out
    none = lambda: _empty(0, like=like)  # noqa: E731  (outputs of a custom op may not alias each other)
    gated, y = _empty(N, HID, like=like), _empty(N, HID, like=like)
    ops.gate_fwd(conv_out, c_act, c_gate, gated)
    # recompute y = Lin_self(gated) + Lin_skip(x_in)
    ops.rowmat_mul(gated, 0, wself_s, 0, y, 0, S, S)
    ops.rowmat_mul(x_in, 0, wskip_s, 0, y, 0, s_in, S, accumulate=True)
    for c in range(3):
        ops.rowmat_mul(gated, S + V * c, wself_v, 0, y, S + V * c, V, V)
        if v_in:
            ops.rowmat_mul(x_in, s_in + v_in * c, wskip_v, 0, y, S + V * c, v_in, V, accumulate=True)
    dy = _empty(N, HID, like=like)
    has_skip, has_scale = skip_w is not None, (s_next is not None and dx_scaled is not None)
    dx_res = _empty(N, HID, like=like) if has_skip else none()
    prod_s = _empty(N, HID, like=like) if has_scale else None
    prod_w = _empty(N, HID, like=like) if has_skip else None
    ops.mix_bwd(None if dx_new is None else dx_new.contiguous(), dx_scaled.contiguous() if has_scale else None, y,
                None if x_res is None else x_res.contiguous(), skip_w, s_next if has_scale else None, dy,
                dx_res if has_skip else None, prod_s, prod_w)
    ds_next = torch.zeros(SO, dtype=torch.float32, device=like.device) if s_next is not None else none()
    dskip_w = _empty(SO, like=like) if has_skip else none()
    if has_scale:
        ops.colsum(prod_s, HID, ds_next, fold_s=S, fold_v=V)
    if has_skip:
        ops.colsum(prod_w, HID, dskip_w, fold_s=S, fold_v=V)
    # the four Linears: dX = dY . W^T, dW = X^T . dY
    dgated, dx_in = _empty(N, HID, like=like), _empty(N, D_in, like=like)
    ops.rowmat_mul(dy, 0, wself_s, 0, dgated, 0, S, S, trans_w=True)
    ops.rowmat_mul(dy, 0, wskip_s, 0, dx_in, 0, S, s_in, trans_w=True)
    dwself_s, dwself_v, dwskip_s = torch.empty_like(wself_s), torch.empty_like(wself_v), torch.empty_like(wskip_s)
    dwskip_v = torch.empty_like(wskip_v) if v_in else none()
    ops.rowmat_dw(gated, 0, dy, 0, dwself_s, 0, S, S)
    ops.rowmat_dw(x_in, 0, dy, 0, dwskip_s, 0, s_in, S)
    for c in range(3):
        ops.rowmat_mul(dy, S + V * c, wself_v, 0, dgated, S + V * c, V, V, trans_w=True)
        ops.rowmat_dw(gated, S + V * c, dy, S + V * c, dwself_v, 0, V, V, accumulate=c > 0)
        if v_in:
            ops.rowmat_mul(dy, S + V * c, wskip_v, 0, dx_in, s_in + v_in * c, V, v_in, trans_w=True)
            ops.rowmat_dw(x_in, s_in + v_in * c, dy, S + V * c, dwskip_v, 0, v_in, V, accumulate=c > 0)
    dconv = _empty(N, GIN, like=like)
    ops.gate_bwd(conv_out, dgated, c_act, c_gate, dconv)
    return [dconv, dx_in, dx_res, dwself_s, dwself_v, dwskip_s, dwskip_v, dskip_w, ds_next]


@block_tail_bwd.register_fake
def _(dx_new, dx_scaled, conv_out, x_in, x_res, wself_s, wself_v, wskip_s, wskip_v, skip_w, s_next, s_in, v_in, c_act, c_gate):
    N = conv_out.shape[0]
    e = conv_out.new_empty(0)
    return [conv_out.new_empty(N, GIN), torch.empty_like(x_in), conv_out.new_empty(N, HID) if skip_w is not None else e,
            torch.empty_like(wself_s), torch.empty_like(wself_v), torch.empty_like(wskip_s),
            torch.empty_like(wskip_v) if wskip_v is not None else e, conv_out.new_empty(SO) if skip_w is not None else e,
            conv_out.new_empty(SO) if s_next is not None else e]


def _tail_setup(ctx, inputs, output):
    conv_out, x_in, x_res, wself_s, wself_v, wskip_s, wskip_v, skip_w, s_next, s_in, v_in, c_act, c_gate = inputs
    ctx.save_for_backward(conv_out, x_in, x_res, wself_s, wself_v, wskip_s, wskip_v, skip_w, s_next)
    ctx.consts = (s_in, v_in, c_act, c_gate)
    ctx.set_materialize_grads(False)


def _tail_backward(ctx, dx_new, dx_scaled):
    conv_out, x_in, x_res, wself_s, wself_v, wskip_s, wskip_v, skip_w, s_next = ctx.saved_tensors
    if s_next is None:
        dx_scaled = None
    if dx_new is None and dx_scaled is None:
        return (None,) * 13
    r = torch.ops.jamun_b200.block_tail_bwd(dx_new, dx_scaled, conv_out, x_in, x_res, wself_s, wself_v, wskip_s, wskip_v, skip_w,
                                            s_next, *ctx.consts)
    opt = lambda t, present: t if present else None  # noqa: E731
    return (r[0], r[1], opt(r[2], x_res is not None and skip_w is not None), r[3], r[4], r[5], opt(r[6], wskip_v is not None),
            opt(r[7], skip_w is not None), opt(r[8], s_next is not None), None, None, None, None)


block_tail.register_autograd(_tail_backward, setup_context=_tail_setup)


# =====================================================================================================================
# output head
# =====================================================================================================================
@torch.library.custom_op(f"{NS_LIB}::head", mutates_args=())
def head(x: Tensor, w1s: Tensor, w1v: Tensor, w2: Tensor, c_gate: float) -> Tensor:
    g = _empty(x.shape[0], 3, like=x)
    if x.shape[0]:
        ops.head(x.contiguous(), w1s.contiguous(), w1v.contiguous(), w2.contiguous(), c_gate, g)
    return g


@head.register_fake
def _(x, w1s, w1v, w2, c_gate):
    return x.new_empty(x.shape[0], 3)


@torch.library.custom_op(f"{NS_LIB}::head_bwd", mutates_args=())
def head_bwd(dg: Tensor, x: Tensor, w1s: Tensor, w1v: Tensor, w2: Tensor, c_gate: float) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    dg, x, w1s, w1v, w2 = dg.contiguous(), x.contiguous(), w1s.contiguous(), w1v.contiguous(), w2.contiguous()
    N = x.shape[0]
    pre, hv = _empty(N, V, like=x), _empty(N, 3 * V, like=x)
    ops.rowmat_mul(x, 0, w1s, S, pre, 0, S, V)
    for c in range(3):
        ops.rowmat_mul(x, S + V * c, w1v, 0, hv, V * c, V, V)
    dpre, dhv, prod = _empty(N, V, like=x), _empty(N, 3 * V, like=x), _empty(N, V, like=x)
    ops.head_bwd(pre, hv, w2, dg, c_gate, dpre, dhv, prod)
    dw2 = _empty(V, like=x)
    ops.colsum(prod, V, dw2)
    dx = _empty(N, HID, like=x)
    ops.rowmat_mul(dpre, 0, w1s, S, dx, 0, V, S, trans_w=True)
    dw1s = torch.zeros_like(w1s)
    ops.rowmat_dw(x, 0, dpre, 0, dw1s, S, S, V)
    dw1v = torch.empty_like(w1v)
    for c in range(3):
        ops.rowmat_mul(dhv, V * c, w1v, 0, dx, S + V * c, V, V, trans_w=True)
        ops.rowmat_dw(x, S + V * c, dhv, V * c, dw1v, 0, V, V, accumulate=c > 0)
    return dx, dw1s, dw1v, dw2


@head_bwd.register_fake
def _(dg, x, w1s, w1v, w2, c_gate):
    return torch.empty_like(x), torch.empty_like(w1s), torch.empty_like(w1v), torch.empty_like(w2)


def _head_setup(ctx, inputs, output):
    x, w1s, w1v, w2, c_gate = inputs
    ctx.save_for_backward(x, w1s, w1v, w2)
    ctx.c_gate = c_gate


def _head_backward(ctx, dg):
    return (*torch.ops.jamun_b200.head_bwd(dg, *ctx.saved_tensors, ctx.c_gate), None)


head.register_autograd(_head_backward, setup_context=_head_setup)


# =====================================================================================================================
# radial MLP hidden layer, atom embedding, noise-conditioning MLP
# =====================================================================================================================
@torch.library.custom_op(f"{NS_LIB}::radial_hidden", mutates_args=())
def radial_hidden(rb: Tensor, ebond: Tensor, rowptr: Tensor, w0r: Tensor, b0eff: Tensor) -> Tensor:
    """h = SiLU(rb . w0r + b0eff[ebond]) for the live edges (rows >= rowptr[N] stay zero)."""
    h = torch.zeros(ebond.numel(), 64, dtype=torch.float32, device=rb.device)
    ops.edge_radial_hidden(rb, ebond, rowptr, w0r.contiguous(), b0eff.contiguous(), h)
    return h


@radial_hidden.register_fake
def _(rb, ebond, rowptr, w0r, b0eff):
    return rb.new_empty(ebond.numel(), 64)


@torch.library.custom_op(f"{NS_LIB}::radial_hidden_bwd", mutates_args=())
def radial_hidden_bwd(dh: Tensor, rb: Tensor, ebond: Tensor, rowptr: Tensor, w0r: Tensor, b0eff: Tensor) -> Tuple[Tensor, Tensor]:
    cap, N = ebond.numel(), rowptr.numel() - 1
    dz = _empty(cap, 64, like=rb)
    ops.radial_bwd(rb, ebond, rowptr, w0r.contiguous(), b0eff.contiguous(), dh.contiguous(), dz)
    e_dev = rowptr[N:]  # live edge count, on the device
    dw0r, db0eff = torch.empty_like(w0r), torch.empty_like(b0eff)
    ops.rowmat_dw(rb, 0, dz, 0, dw0r, 0, 32, 64, rows=cap, rows_dev=e_dev)
    for flag in (0, 1):
        ops.colsum(dz, 64, db0eff, flag=ebond, flag_value=flag, rows=cap, rows_dev=e_dev, ocol=64 * flag)
    return dw0r, db0eff


@radial_hidden_bwd.register_fake
def _(dh, rb, ebond, rowptr, w0r, b0eff):
    return torch.empty_like(w0r), torch.empty_like(b0eff)


def _radial_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _radial_backward(ctx, dh):
    rb, ebond, rowptr, w0r, b0eff = ctx.saved_tensors
    dw0r, db0eff = torch.ops.jamun_b200.radial_hidden_bwd(dh, rb, ebond, rowptr, w0r, b0eff)
    return None, None, None, dw0r, db0eff


radial_hidden.register_autograd(_radial_backward, setup_context=_radial_setup)


@torch.library.custom_op(f"{NS_LIB}::atom_embed", mutates_args=())
def atom_embed(idx0: Tensor, idx1: Tensor, idx2: Tensor, idx3: Optional[Tensor], tab0: Tensor, tab1: Tensor, tab2: Tensor,
               tab3: Tensor, scale: Tensor) -> Tensor:
    tabs = [t.contiguous() for t in (tab0, tab1, tab2, tab3)]
    return ops.atom_embed([idx0, idx1, idx2, idx3], tabs, scale.contiguous())


@atom_embed.register_fake
def _(idx0, idx1, idx2, idx3, tab0, tab1, tab2, tab3, scale):
    return tab0.new_empty(idx0.numel(), tab0.shape[1] + tab1.shape[1] + tab2.shape[1] + tab3.shape[1])


@torch.library.custom_op(f"{NS_LIB}::atom_embed_bwd", mutates_args=())
def atom_embed_bwd(dx0: Tensor, idx0: Tensor, idx1: Tensor, idx2: Tensor, idx3: Optional[Tensor], tab0: Tensor, tab1: Tensor,
                   tab2: Tensor, tab3: Tensor, scale: Tensor) -> List[Tensor]:
    tabs = [t.contiguous() for t in (tab0, tab1, tab2, tab3)]
    dtabs = [torch.empty_like(t) for t in tabs]
    dx0 = dx0.contiguous()
    prod = torch.empty_like(dx0)
    ops.embed_bwd([idx0, idx1, idx2, idx3], tabs, scale.contiguous(), dx0, dtabs, prod)
    dscale = torch.empty_like(scale)
    ops.colsum(prod, prod.shape[1], dscale)
    return dtabs + [dscale]


@atom_embed_bwd.register_fake
def _(dx0, idx0, idx1, idx2, idx3, tab0, tab1, tab2, tab3, scale):
    return [torch.empty_like(t) for t in (tab0, tab1, tab2, tab3, scale)]


def _embed_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _embed_backward(ctx, dx0):
    r = torch.ops.jamun_b200.atom_embed_bwd(dx0, *ctx.saved_tensors)
    return (None, None, None, None, *r)


atom_embed.register_autograd(_embed_backward, setup_context=_embed_setup)


@torch.library.custom_op(f"{NS_LIB}::noise_mlp", mutates_args=())
def noise_mlp(w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, c_noise: float, apply_sigmoid: bool) -> Tensor:
    return ops.noise_mlp(w1.contiguous(), b1.contiguous(), w2.contiguous(), b2.contiguous(), c_noise, apply_sigmoid)


@noise_mlp.register_fake
def _(w1, b1, w2, b2, c_noise, apply_sigmoid):
    return torch.empty_like(b1)


@torch.library.custom_op(f"{NS_LIB}::noise_mlp_bwd", mutates_args=())
def noise_mlp_bwd(dout: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, c_noise: float, apply_sigmoid: bool
                  ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    w1, b1, w2, b2 = w1.contiguous(), b1.contiguous(), w2.contiguous(), b2.contiguous()
    grads = [torch.empty_like(t) for t in (w1, b1, w2, b2)]
    ops.noise_mlp_bwd(w1, b1, w2, b2, c_noise, apply_sigmoid, dout.contiguous(), *grads)
    return tuple(grads)


@noise_mlp_bwd.register_fake
def _(dout, w1, b1, w2, b2, c_noise, apply_sigmoid):
    return tuple(torch.empty_like(t) for t in (w1, b1, w2, b2))


def _nmlp_setup(ctx, inputs, output):
    w1, b1, w2, b2, c_noise, apply_sigmoid = inputs
    ctx.save_for_backward(w1, b1, w2, b2)
    ctx.consts = (c_noise, apply_sigmoid)


def _nmlp_backward(ctx, dout):
    return (*torch.ops.jamun_b200.noise_mlp_bwd(dout, *ctx.saved_tensors, *ctx.consts), None, None)


noise_mlp.register_autograd(_nmlp_backward, setup_context=_nmlp_setup)


# =====================================================================================================================
# denoising loss and Kabsch alignment
# =====================================================================================================================
@torch.library.custom_op(f"{NS_LIB}::combine_xhat", mutates_args=())
def combine_xhat(g: Tensor, ybar: Tensor, chain_ptr: Tensor, c_skip: float, c_out: float, center: bool) -> Tensor:
    """xhat = center_chain(c_skip*ybar + c_out*g) (model/denoiser.py:200,213-215)."""
    return ops.combine_xhat(g.contiguous(), ybar.contiguous(), chain_ptr, c_skip, c_out, center, torch.empty_like(g))


@combine_xhat.register_fake
def _(g, ybar, chain_ptr, c_skip, c_out, center):
    return torch.empty_like(g)


@torch.library.custom_op(f"{NS_LIB}::combine_xhat_bwd", mutates_args=())
def combine_xhat_bwd(dxhat: Tensor, chain_ptr: Tensor, c_out: float, center: bool) -> Tensor:
    return ops.combine_xhat(dxhat.contiguous(), None, chain_ptr, 0.0, c_out, center, torch.empty_like(dxhat))


@combine_xhat_bwd.register_fake
def _(dxhat, chain_ptr, c_out, center):
    return torch.empty_like(dxhat)


def _comb_setup(ctx, inputs, output):
    g, ybar, chain_ptr, c_skip, c_out, center = inputs
    ctx.save_for_backward(chain_ptr)
    ctx.consts = (c_out, center)


def _comb_backward(ctx, dxhat):
    (chain_ptr,) = ctx.saved_tensors
    return torch.ops.jamun_b200.combine_xhat_bwd(dxhat, chain_ptr, *ctx.consts), None, None, None, None, None


combine_xhat.register_autograd(_comb_backward, setup_context=_comb_setup)


@torch.library.custom_op(f"{NS_LIB}::coordinate_loss", mutates_args=())
def coordinate_loss(xhat: Tensor, x: Tensor, chain_of: Tensor, chain_ptr: Tensor, loss_weight: Optional[Tensor], scale: float,
                    sigma: float) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (loss [G], raw_coordinate_loss [G], scaled_rmsd [G]) (model/denoiser.py:251-287)."""
    G = chain_ptr.numel() - 1
    loss, raw, rmsd = (_empty(G, like=xhat) for _ in range(3))
    ops.loss_fwd(xhat.contiguous(), x.contiguous(), chain_ptr, scale, sigma, loss_weight, loss, raw, rmsd)
    return loss, raw, rmsd


@coordinate_loss.register_fake
def _(xhat, x, chain_of, chain_ptr, loss_weight, scale, sigma):
    G = chain_ptr.numel() - 1
    return xhat.new_empty(G), xhat.new_empty(G), xhat.new_empty(G)


@torch.library.custom_op(f"{NS_LIB}::coordinate_loss_bwd", mutates_args=())
def coordinate_loss_bwd(dloss: Tensor, xhat: Tensor, x: Tensor, chain_of: Tensor, chain_ptr: Tensor, loss_weight: Optional[Tensor],
                        scale: float) -> Tensor:
    dxhat = torch.empty_like(xhat)
    ops.loss_bwd(xhat.contiguous(), x.contiguous(), chain_of, chain_ptr, scale, loss_weight, dloss.contiguous(), dxhat)
    return dxhat


@coordinate_loss_bwd.register_fake
def _(dloss, xhat, x, chain_of, chain_ptr, loss_weight, scale):
    return torch.empty_like(xhat)


def _loss_setup(ctx, inputs, output):
    xhat, x, chain_of, chain_ptr, loss_weight, scale, sigma = inputs
    ctx.save_for_backward(xhat, x, chain_of, chain_ptr, loss_weight)
    ctx.scale = scale
    ctx.set_materialize_grads(False)


def _loss_backward(ctx, dloss, draw, drmsd):
    xhat, x, chain_of, chain_ptr, loss_weight = ctx.saved_tensors
    if dloss is None:
        return (None,) * 7
    # only the loss carries a gradient (the reference back-propagates `loss` and logs the other two, denoiser.py:305-319)
    return torch.ops.jamun_b200.coordinate_loss_bwd(dloss, xhat, x, chain_of, chain_ptr, loss_weight, ctx.scale), None, None, None, \
        None, None, None


coordinate_loss.register_autograd(_loss_backward, setup_context=_loss_setup)


@torch.library.custom_op(f"{NS_LIB}::kabsch_align", mutates_args=())
def kabsch_align(y: Tensor, x: Tensor, chain_ptr: Tensor) -> Tensor:
    """align_A_to_B_batched (utils/align.py:123-126): per chain, y -> R y + t minimising |R y + t - x|."""
    out = torch.empty_like(y)
    ops.kabsch_align(y.contiguous(), x.contiguous(), chain_ptr, out)
    return out


@kabsch_align.register_fake
def _(y, x, chain_ptr):
    return torch.empty_like(y)
