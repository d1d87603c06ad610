"""Torch-tensor front end of the C ABI: argument checking, stream plumbing, nothing else.

Every function enqueues hand-written sm_100a kernels on torch's current CUDA stream (so the calls
are CUDA-graph capturable) and returns torch tensors that alias caller-owned or freshly allocated
device memory.  CPU tensors are rejected: there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib

S, V, HID, GATE_IN, S0, EDGE_HID, NBASIS = 120, 32, 216, 248, 56, 64, 32

LAUNCHES = 0  # number of kernel launches issued through this module (bench.py's gpu_launches claim)

Tensor = torch.Tensor
_T = torch.ops.jamun_b200


def _op(name: str, mutates=()):
    """Register a launcher as the torch.library operator jamun_b200::<name> (thin custom-op layer over the C ABI).
    Launchers write into caller-owned buffers (declared in `mutates`) and return nothing, so their fake implementation is
    empty; operators that allocate their result are registered with an explicit fake below."""
    def deco(fn):
        op = torch.library.custom_op(f"jamun_b200::{name}", fn, mutates_args=tuple(mutates), device_types="cuda")
        op.register_fake(lambda *a, **k: None)
        return op
    return deco


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _ptr(t: Optional[torch.Tensor], dtype=torch.float32) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("jamun_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (what torch.cuda.current_stream().cuda_stream returns, without
    the ~15 us of Python-side device bookkeeping that call costs: it is made once per kernel launch)."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


@_op("noise_mlp_", mutates=("out",))
def _noise_mlp(w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, c_noise: float, apply_sigmoid: bool, out: Tensor) -> None:
    n = b1.numel()
    rc = _lib.lib().jamun_noise_mlp(_ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), float(c_noise), n, int(apply_sigmoid),
                                    _ptr(out), _stream())
    _lib.check(rc, "jamun_noise_mlp")
    _count()


def noise_mlp(w1, b1, w2, b2, c_noise: float, apply_sigmoid: bool, out=None):
    out = torch.empty(b1.numel(), device=b1.device, dtype=torch.float32) if out is None else out
    _T.noise_mlp_(w1, b1, w2, b2, float(c_noise), bool(apply_sigmoid), out)
    return out


@_op("atom_embed_", mutates=("out",))
def _atom_embed(idx: List[Optional[Tensor]], tabs: List[Tensor], scale: Optional[Tensor], out: Tensor) -> None:
    N = idx[0].numel()
    dims = [t.shape[1] for t in tabs]
    i32 = torch.int32
    rc = _lib.lib().jamun_atom_embed(_ptr(idx[0], i32), _ptr(idx[1], i32), _ptr(idx[2], i32),
                                     _ptr(idx[3], i32) if idx[3] is not None else None,
                                     _ptr(tabs[0]), _ptr(tabs[1]), _ptr(tabs[2]), _ptr(tabs[3]), *dims,
                                     _ptr(scale), N, _ptr(out), _stream())
    _lib.check(rc, "jamun_atom_embed")
    _count()


def atom_embed(idx, tabs, scale, out=None):
    out = torch.empty(idx[0].numel(), sum(t.shape[1] for t in tabs), device=tabs[0].device, dtype=torch.float32) if out is None else out
    _T.atom_embed_(list(idx), list(tabs), scale, out)
    return out


@_op("center_scale_", mutates=("ybar", "p"))
def _center_scale(y: Tensor, chain_ptr: Tensor, c_in: float, ybar: Tensor, p: Tensor, center: bool) -> None:
    G = chain_ptr.numel() - 1
    rc = _lib.lib().jamun_center_scale(_ptr(y), _ptr(chain_ptr, torch.int32), G, int(center), float(c_in), _ptr(ybar), _ptr(p),
                                       _stream())
    _lib.check(rc, "jamun_center_scale")
    _count()


def center_scale(y, chain_ptr, c_in: float, ybar=None, p=None, center: bool = True):
    ybar = torch.empty_like(y) if ybar is None else ybar
    p = torch.empty_like(y) if p is None else p
    _T.center_scale_(y, chain_ptr, float(c_in), ybar, p, bool(center))
    return ybar, p


@_op("radius_csr", mutates=("scratch", "rowptr", "col", "edst", "ebond"))
def _radius_csr(pos: Tensor, chain_of: Tensor, chain_ptr: Tensor, r2: float, max_num_neighbors: int, bond_rowptr: Tensor,
                bond_src: Tensor, scratch: Tensor, rowptr: Tensor, col: Tensor, edst: Tensor, ebond: Tensor) -> None:
    N = pos.shape[0]
    i32 = torch.int32
    rc = _lib.lib().jamun_radius_csr(_ptr(pos), _ptr(chain_of, i32), _ptr(chain_ptr, i32), N, float(r2),
                                     int(max_num_neighbors), _ptr(bond_rowptr, i32), _ptr(bond_src, i32),
                                     _ptr(scratch, i32), _ptr(rowptr, i32), _ptr(col, i32), _ptr(edst, i32),
                                     _ptr(ebond, torch.uint8), _stream())
    _lib.check(rc, "jamun_radius_csr")
    _count(3)


def radius_csr(pos, chain_of, chain_ptr, r2: float, max_num_neighbors: int, bond_rowptr, bond_src, scratch, rowptr, col, edst, ebond):
    _T.radius_csr(pos, chain_of, chain_ptr, float(r2), int(max_num_neighbors), bond_rowptr, bond_src, scratch, rowptr, col, edst, ebond)


@_op("radius_csr_cells", mutates=("scratch", "nbr", "rowptr", "col", "edst", "ebond"))
def _radius_csr_cells(pos: Tensor, chain_ptr: Tensor, max_chain: int, r2: float, r_cut: float, max_num_neighbors: int,
                      bond_rowptr: Tensor, bond_src: Tensor, scratch: Tensor, nbr: Tensor, rowptr: Tensor, col: Tensor, edst: Tensor,
                      ebond: Tensor) -> None:
    N, G = pos.shape[0], chain_ptr.numel() - 1
    i32 = torch.int32
    rc = _lib.lib().jamun_radius_csr_cells(_ptr(pos), _ptr(chain_ptr, i32), G, N, int(max_chain), float(r2), float(r_cut),
                                           int(max_num_neighbors), _ptr(bond_rowptr, i32), _ptr(bond_src, i32), _ptr(scratch, i32),
                                           _ptr(nbr, i32), _ptr(rowptr, i32), _ptr(col, i32), _ptr(edst, i32),
                                           _ptr(ebond, torch.uint8), _stream())
    _lib.check(rc, "jamun_radius_csr_cells")
    _count(3)


def radius_csr_cells(pos, chain_ptr, max_chain: int, r2: float, r_cut: float, max_num_neighbors: int, bond_rowptr, bond_src, scratch,
                     nbr, rowptr, col, edst, ebond):
    _T.radius_csr_cells(pos, chain_ptr, int(max_chain), float(r2), float(r_cut), int(max_num_neighbors), bond_rowptr, bond_src, scratch,
                        nbr, rowptr, col, edst, ebond)


@_op("edge_geom", mutates=("rhat", "rb"))
def _edge_geom(p: Tensor, rowptr: Tensor, col: Tensor, edst: Tensor, mu: Tensor, step: float, rhat: Tensor, rb: Tensor) -> None:
    N, cap = p.shape[0], col.numel()
    i32 = torch.int32
    rc = _lib.lib().jamun_edge_geom(_ptr(p), _ptr(rowptr, i32), _ptr(col, i32), _ptr(edst, i32), N, cap, _ptr(mu),
                                    float(step), _ptr(rhat), _ptr(rb), _stream())
    _lib.check(rc, "jamun_edge_geom")
    _count()


def edge_geom(p, rowptr, col, edst, mu, step: float, rhat, rb):
    _T.edge_geom(p, rowptr, col, edst, mu, float(step), rhat, rb)


@_op("edge_radial_hidden", mutates=("h",))
def _edge_radial_hidden(rb: Tensor, ebond: Tensor, rowptr: Tensor, w0r: Tensor, b0eff: Tensor, h: Tensor) -> None:
    N, cap = rowptr.numel() - 1, ebond.numel()
    rc = _lib.lib().jamun_edge_radial_hidden(_ptr(rb), _ptr(ebond, torch.uint8), _ptr(rowptr, torch.int32), N, cap,
                                             _ptr(w0r), _ptr(b0eff), _ptr(h), _stream())
    _lib.check(rc, "jamun_edge_radial_hidden")
    _count()


def edge_radial_hidden(rb, ebond, rowptr, w0r, b0eff, h):
    _T.edge_radial_hidden(rb, ebond, rowptr, w0r, b0eff, h)


@_op("edge_radial_hidden_all", mutates=("h_all",))
def _edge_radial_hidden_all(rb: Tensor, ebond: Tensor, rowptr: Tensor, w0r_all: Tensor, b0eff_all: Tensor, h_all: Tensor) -> None:
    N, cap, layers = rowptr.numel() - 1, ebond.numel(), w0r_all.shape[0]
    assert h_all.shape[0] == layers and h_all.shape[1] == cap
    rc = _lib.lib().jamun_edge_radial_hidden_all(_ptr(rb), _ptr(ebond, torch.uint8), _ptr(rowptr, torch.int32), N, cap,
                                                 _ptr(w0r_all), _ptr(b0eff_all), layers, _ptr(h_all), _stream())
    _lib.check(rc, "jamun_edge_radial_hidden_all")
    _count()


def edge_radial_hidden_all(rb, ebond, rowptr, w0r_all, b0eff_all, h_all):
    _T.edge_radial_hidden_all(rb, ebond, rowptr, w0r_all, b0eff_all, h_all)


def radial_pack_frag(w0r_all: Tensor) -> Tensor:
    """Pre-split, fragment-ordered weight images of the tensor-core radial hidden kernel ([layers * 4096] floats; plan build)."""
    layers = w0r_all.shape[0]
    img = torch.empty(layers * 4096, dtype=torch.float32, device=w0r_all.device)
    rc = _lib.lib().jamun_radial_pack_frag(_ptr(w0r_all), layers, _ptr(img), _stream())
    _lib.check(rc, "jamun_radial_pack_frag")
    _count()
    return img


@_op("edge_radial_hidden_mma", mutates=("h_all",))
def _edge_radial_hidden_mma(rb: Tensor, ebond: Tensor, rowptr: Tensor, img: Tensor, b0eff_all: Tensor, h_all: Tensor) -> None:
    N, cap, layers = rowptr.numel() - 1, ebond.numel(), b0eff_all.shape[0]
    assert h_all.shape[0] == layers and h_all.shape[1] == cap and img.numel() == layers * 4096
    rc = _lib.lib().jamun_edge_radial_hidden_mma(_ptr(rb), _ptr(ebond, torch.uint8), _ptr(rowptr, torch.int32), N, cap, _ptr(img),
                                                 _ptr(b0eff_all), layers, _ptr(h_all), _stream())
    _lib.check(rc, "jamun_edge_radial_hidden_mma")
    _count()


def edge_radial_hidden_mma(rb, ebond, rowptr, img, b0eff_all, h_all):
    _T.edge_radial_hidden_mma(rb, ebond, rowptr, img, b0eff_all, h_all)


@_op("conv_fwd_simt", mutates=("out",))
def _conv_fwd(x: Tensor, s_in: int, v_in: int, rowptr: Tensor, col: Tensor, h: Tensor, rhat: Tensor, m0: Tensor, m1: Tensor,
              alpha0: float, alpha1: float, out: Tensor) -> None:
    N = x.shape[0]
    assert x.shape[1] == s_in + 3 * v_in and out.shape == (N, GATE_IN)
    i32 = torch.int32
    rc = _lib.lib().jamun_conv_fwd(_ptr(x), s_in, v_in, _ptr(rowptr, i32), _ptr(col, i32), _ptr(h), _ptr(rhat),
                                   _ptr(m0), _ptr(m1), float(alpha0), float(alpha1), N, _ptr(out), _stream())
    _lib.check(rc, "jamun_conv_fwd")
    _count()


def conv_fwd(x, s_in: int, v_in: int, rowptr, col, h, rhat, m0, m1, alpha0: float, alpha1: float, out):
    _T.conv_fwd_simt(x, int(s_in), int(v_in), rowptr, col, h, rhat, m0, m1, float(alpha0), float(alpha1), out)
    return out


def conv_build_a(x, s_in: int, v_in: int, rowptr, col, h, rhat, y, max_degree: int, row0: int, nrows: int, rows_pad: int, a0_ptr: int,
                 a1_ptr: int, a1_comp_stride: int, p2_ptr: int, p2_ld: int, p2_scale: float, inv_deg):
    i32 = torch.int32
    rc = _lib.lib().jamun_conv_build_a(_ptr(x), s_in, v_in, _ptr(rowptr, i32), _ptr(col, i32), _ptr(h), _ptr(rhat), _ptr(y),
                                       int(max_degree), row0, nrows, rows_pad, a0_ptr, a1_ptr, int(a1_comp_stride), p2_ptr, p2_ld,
                                       float(p2_scale), _ptr(inv_deg), _stream())
    _lib.check(rc, "jamun_conv_build_a")
    _count()


def conv_build_tc(x, s_in: int, v_in: int, rowptr, col, h, rhat, row0: int, nrows: int, rows_pad: int, a0_ptr: int, a1_ptr: int,
                  a1_comp_stride: int, inv_deg=None, tiled: bool = False):
    """tiled: write the tile-major operand layout [row / 128][stage][128][32] (gemm_f16x3(a_tile_major=True))."""
    i32 = torch.int32
    fn = _lib.lib().jamun_conv_build_tc_tiled if tiled else _lib.lib().jamun_conv_build_tc
    rc = fn(_ptr(x), s_in, v_in, _ptr(rowptr, i32), _ptr(col, i32), _ptr(h), _ptr(rhat), row0, nrows, rows_pad, a0_ptr, a1_ptr,
            int(a1_comp_stride), _ptr(inv_deg), _stream())
    _lib.check(rc, "jamun_conv_build_tc")
    _count()


def conv_p2(rowptr, src_rowptr, src_eid, h, rhat, y, t_edge, p2_ptr: int, p2_ld: int, p2_scale: float, inv_deg=None):
    i32 = torch.int32
    N = rowptr.numel() - 1
    rc = _lib.lib().jamun_conv_p2(_ptr(rowptr, i32), _ptr(src_rowptr, i32), _ptr(src_eid, i32), _ptr(h), _ptr(rhat), _ptr(y), N,
                                  _ptr(t_edge), p2_ptr, p2_ld, float(p2_scale), _ptr(inv_deg), _stream())
    _lib.check(rc, "jamun_conv_p2")
    _count(2 if p2_ptr else 1)


@_op("csr_by_source", mutates=("scratch", "src_rowptr", "src_eid"))
def _csr_by_source(rowptr: Tensor, col: Tensor, scratch: Tensor, src_rowptr: Tensor, src_eid: Tensor) -> None:
    i32 = torch.int32
    rc = _lib.lib().jamun_csr_by_source(_ptr(rowptr, i32), _ptr(col, i32), rowptr.numel() - 1, col.numel(), _ptr(scratch, i32),
                                        _ptr(src_rowptr, i32), _ptr(src_eid, i32), _stream())
    _lib.check(rc, "jamun_csr_by_source")
    _count(4)


def csr_by_source(rowptr, col, scratch, src_rowptr, src_eid):
    _T.csr_by_source(rowptr, col, scratch, src_rowptr, src_eid)


@_op("pack_rows", mutates=("a",))
def _pack_rows(x: Tensor, col0: int, ncols: int, rows_pad: int, a: Tensor) -> None:
    rc = _lib.lib().jamun_pack_rows(_ptr(x), x.shape[1], col0, ncols, x.shape[0], rows_pad, _ptr(a), _stream())
    _lib.check(rc, "jamun_pack_rows")
    _count()


def pack_rows(x, col0: int, ncols: int, rows_pad: int, a):
    _T.pack_rows(x, int(col0), int(ncols), int(rows_pad), a)


def pack_b(src, n_stages: int, n_pad: int, row_map=None, k_src: Optional[int] = None, n_valid: Optional[int] = None,
           n_inner: Optional[int] = None, outer_rows: int = 0, col_blocks: int = 1, transpose: bool = False, out=None):
    """Row-major weights -> (hi | lo) stage images of the GEMM's B operand ([col_blocks, n_stages, 2, n_pad*32])."""
    assert src.dim() == 2
    k_src = src.shape[0] if k_src is None else k_src
    n_valid = src.shape[1] if n_valid is None else n_valid
    n_inner = max(n_valid, 1) if n_inner is None else n_inner
    shape = (col_blocks, n_stages, 2, n_pad * 32) if col_blocks > 1 else (n_stages, 2, n_pad * 32)
    out = torch.empty(shape, dtype=torch.float32, device=src.device) if out is None else out
    rc = _lib.lib().jamun_pack_b(_ptr(src), src.stride(0), _ptr(row_map, torch.int32), k_src, n_stages, n_valid, n_inner,
                                 outer_rows, n_pad, col_blocks, int(transpose), _ptr(out), _stream())
    _lib.check(rc, "jamun_pack_b")
    _count()
    return out


def f16_scale(*weights) -> float:
    """Power-of-two pre-scale for fp16-split weight images: the largest magnitude lands in [2^12, 2^13) (fp16 max 65504), so
    every element within 2^-16 of it keeps a normal-range remainder.  Host sync (amax); call when a plan is built."""
    import math

    amax = max(float(w.detach().abs().max()) for w in weights if w is not None and w.numel())
    if not math.isfinite(amax) or amax <= 0.0:
        return 1.0
    return float(2.0 ** (12 - math.floor(math.log2(amax))))


def pack_b_f16(src, n_stages: int, n_pad: int, scale: float, row_map=None, k_src: Optional[int] = None, n_valid: Optional[int] = None,
               n_inner: Optional[int] = None, outer_rows: int = 0, col_blocks: int = 1, transpose: bool = False, out=None):
    """Row-major weights -> fp16-split stage images of jamun_gemm_f16x3 ([col_blocks, n_stages, n_pad*32] 4-byte words)."""
    assert src.dim() == 2
    k_src = src.shape[0] if k_src is None else k_src
    n_valid = src.shape[1] if n_valid is None else n_valid
    n_inner = max(n_valid, 1) if n_inner is None else n_inner
    shape = (col_blocks, n_stages, n_pad * 32) if col_blocks > 1 else (n_stages, n_pad * 32)
    out = torch.empty(shape, dtype=torch.float32, device=src.device) if out is None else out
    rc = _lib.lib().jamun_pack_b_f16(_ptr(src), src.stride(0), _ptr(row_map, torch.int32), k_src, n_stages, n_valid, n_inner,
                                     outer_rows, n_pad, col_blocks, int(transpose), float(scale), _ptr(out), _stream())
    _lib.check(rc, "jamun_pack_b_f16")
    _count()
    return out


def gemm_f16x3(a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, rows: int, rows_pad: int, row_scale_ptr, out_ptr,
               out_ld: int, addend_ptrs=None, addend_ld=None, col_blocks: int = 1, b_block_floats: int = 0, k_splits: int = 1,
               partial=None, status=None, addend_scale=None, a_tile_major: bool = False):
    """fp16-split form of gemm_tf32x3 (B images from pack_b_f16, alpha already divided by their scale).  status: int32 device
    word whose bit 0 reports an A value outside the fp16 range."""
    n = len(a_ptrs)
    VP, IA, FA = C.c_void_p * n, C.c_int * n, C.c_float * n
    ad = VP(*addend_ptrs) if addend_ptrs is not None else None
    adl = IA(*addend_ld) if addend_ld is not None else None
    if k_splits > 1:
        assert partial is not None and partial.numel() >= k_splits * rows * out_ld
    ads = FA(*addend_scale) if addend_scale is not None else None
    rc = _lib.lib().jamun_gemm_f16x3(n, VP(*a_ptrs), VP(*b_ptrs), IA(*n_stages), IA(*n_pad), IA(*n_valid), IA(*out_col), FA(*alpha),
                                     ad, adl, ads, col_blocks, b_block_floats, rows, rows_pad, row_scale_ptr, out_ptr, out_ld,
                                     int(k_splits), _ptr(partial), _ptr(status, torch.int32), int(a_tile_major), _stream())
    _lib.check(rc, "jamun_gemm_f16x3")
    _count(2 if k_splits > 1 else 1)


def gemm_f16x3_fused(a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, rows: int, rows_pad: int, row_scale_ptr, epi: dict,
                     addend_ptrs=None, addend_ld=None, addend_scale=None, status=None, a_tile_major: bool = False):
    """gemm_f16x3 with a fused ConvBlock epilogue (jamun_gemm_epilogue): epi = dict(mode=1|2, op_s=ptr, op_v=ptr,
    op_v_comp_stride=floats, op_rows_pad=rows, c_act=, c_gate= | x_res=ptr, skip_w=ptr, s_next=ptr, x_new=ptr, x_scaled=ptr)."""
    n = len(a_ptrs)
    VP, IA, FA = C.c_void_p * n, C.c_int * n, C.c_float * n
    ad = VP(*addend_ptrs) if addend_ptrs is not None else None
    adl = IA(*addend_ld) if addend_ld is not None else None
    ads = FA(*addend_scale) if addend_scale is not None else None
    e = _lib.GemmEpilogue(**{k: (v if v is not None else None) for k, v in epi.items()})
    rc = _lib.lib().jamun_gemm_f16x3_fused(n, VP(*a_ptrs), VP(*b_ptrs), IA(*n_stages), IA(*n_pad), IA(*n_valid), IA(*out_col),
                                           FA(*alpha), ad, adl, ads, rows, rows_pad, row_scale_ptr, _ptr(status, torch.int32),
                                           int(a_tile_major), C.byref(e), _stream())
    _lib.check(rc, "jamun_gemm_f16x3_fused")
    _count()


def gemm_tf32x3(a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, rows: int, rows_pad: int, row_scale_ptr,
                out_ptr, out_ld: int, addend_ptrs=None, addend_ld=None, col_blocks: int = 1, b_block_floats: int = 0):
    """Raw-pointer front end (segments are slices of larger workspaces).  All lists have one entry per segment."""
    n = len(a_ptrs)
    VP, IA, FA = C.c_void_p * n, C.c_int * n, C.c_float * n
    ad = VP(*addend_ptrs) if addend_ptrs is not None else None
    adl = IA(*addend_ld) if addend_ld is not None else None
    rc = _lib.lib().jamun_gemm_tf32x3(n, VP(*a_ptrs), VP(*b_ptrs), IA(*n_stages), IA(*n_pad), IA(*n_valid), IA(*out_col),
                                      FA(*alpha), ad, adl, col_blocks, b_block_floats, rows, rows_pad, row_scale_ptr, out_ptr,
                                      out_ld, _stream())
    _lib.check(rc, "jamun_gemm_tf32x3")
    _count()


def gemm_tf32x3_splitk(a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, rows: int, rows_pad: int, row_scale_ptr, out_ptr,
                       out_ld: int, k_splits: int, partial):
    """Split-K front end: partial is a [k_splits, rows, out_ld] fp32 scratch tensor."""
    n = len(a_ptrs)
    VP, IA, FA = C.c_void_p * n, C.c_int * n, C.c_float * n
    assert partial.numel() >= k_splits * rows * out_ld
    rc = _lib.lib().jamun_gemm_tf32x3_splitk(n, VP(*a_ptrs), VP(*b_ptrs), IA(*n_stages), IA(*n_pad), IA(*n_valid), IA(*out_col),
                                             FA(*alpha), rows, rows_pad, row_scale_ptr, out_ptr, out_ld, int(k_splits),
                                             _ptr(partial), _stream())
    _lib.check(rc, "jamun_gemm_tf32x3_splitk")
    _count(2 if k_splits > 1 else 1)


@_op("block_tail_", mutates=("x_new", "x_scaled"))
def _block_tail(conv: Tensor, x_in: Tensor, s_in: int, v_in: int, x_res: Optional[Tensor], wself_s: Tensor, wself_v: Tensor,
                wskip_s: Tensor, wskip_v: Optional[Tensor], skip_w: Optional[Tensor], s_next: Optional[Tensor], c_act: float,
                c_gate: float, x_new: Tensor, x_scaled: Optional[Tensor], vadd: Optional[Tensor]) -> None:
    N = conv.shape[0]
    rc = _lib.lib().jamun_block_tail(_ptr(conv), _ptr(vadd), _ptr(x_in), s_in, v_in, _ptr(x_res), _ptr(wself_s), _ptr(wself_v),
                                     _ptr(wskip_s), _ptr(wskip_v), _ptr(skip_w), _ptr(s_next), float(c_act),
                                     float(c_gate), N, _ptr(x_new), _ptr(x_scaled), _stream())
    _lib.check(rc, "jamun_block_tail")
    _count()


def block_tail(conv, x_in, s_in: int, v_in: int, x_res, wself_s, wself_v, wskip_s, wskip_v, skip_w, s_next,
               c_act: float, c_gate: float, x_new, x_scaled, vadd=None):
    _T.block_tail_(conv, x_in, int(s_in), int(v_in), x_res, wself_s, wself_v, wskip_s, wskip_v, skip_w, s_next, float(c_act),
                   float(c_gate), x_new, x_scaled, vadd)


def tail_pack(conv, vadd, x_in, s_in: int, v_in: int, c_act: float, c_gate: float, rows_pad: int, a_s_ptr: int, a_v_ptr: int,
              a_v_comp_stride: int, rowptr=None, rhat=None, t_edge=None, p2_scale: float = 0.0, conv_has_v: bool = True):
    rc = _lib.lib().jamun_tail_pack(_ptr(conv), _ptr(vadd), _ptr(x_in), s_in, v_in, float(c_act), float(c_gate), conv.shape[0],
                                    rows_pad, a_s_ptr, a_v_ptr, int(a_v_comp_stride), _ptr(rowptr, torch.int32), _ptr(rhat),
                                    _ptr(t_edge), float(p2_scale), int(conv_has_v), _stream())
    _lib.check(rc, "jamun_tail_pack")
    _count()


@_op("tail_mix", mutates=("x_new", "x_scaled", "xs_op"))
def _tail_mix(y: Tensor, x_res: Optional[Tensor], skip_w: Optional[Tensor], s_next: Optional[Tensor], x_new: Tensor,
              x_scaled: Optional[Tensor], xs_op: Optional[Tensor], rows_pad: int) -> None:
    rc = _lib.lib().jamun_tail_mix(_ptr(y), _ptr(x_res), _ptr(skip_w), _ptr(s_next), y.shape[0], _ptr(x_new), _ptr(x_scaled),
                                   _ptr(xs_op), rows_pad, _stream())
    _lib.check(rc, "jamun_tail_mix")
    _count()


def tail_mix(y, x_res, skip_w, s_next, x_new, x_scaled, xs_op=None, rows_pad: int = 0):
    _T.tail_mix(y, x_res, skip_w, s_next, x_new, x_scaled, xs_op, int(rows_pad))


@_op("head_", mutates=("g",))
def _head(x: Tensor, w1_s: Tensor, w1_v: Tensor, w2: Tensor, c_gate: float, g: Tensor) -> None:
    rc = _lib.lib().jamun_head(_ptr(x), _ptr(w1_s), _ptr(w1_v), _ptr(w2), float(c_gate), x.shape[0], _ptr(g), _stream())
    _lib.check(rc, "jamun_head")
    _count()


def head(x, w1_s, w1_v, w2, c_gate: float, g):
    _T.head_(x, w1_s, w1_v, w2, float(c_gate), g)
    return g


_WP_FLOATS = ("c_in", "c_skip", "c_out", "sigma2", "delta", "u", "a", "z_sqrt_u", "beta", "clip")
_WP_INTS = ("first", "last", "center", "seed", "step")
_U64 = (1 << 64) - 1


def _signed64(v: int) -> int:  # operator arguments are int64: 64-bit seeds travel as two's complement
    v &= _U64
    return v - (1 << 64) if v >= (1 << 63) else v


@_op("baoab_step", mutates=("y", "v", "ybar", "p", "xhat", "score", "traj_y", "traj_xhat", "traj_score"))
def _walk_step(y: Tensor, v: Tensor, ybar: Tensor, p: Tensor, g: Optional[Tensor], chain_ptr: Tensor, fparams: List[float],
               iparams: List[int], noise: Optional[Tensor], xhat: Optional[Tensor], score: Optional[Tensor],
               traj_y: Optional[Tensor], traj_xhat: Optional[Tensor], traj_score: Optional[Tensor], score_in: Optional[Tensor],
               dev_state: Optional[Tensor]) -> None:
    prm = _lib.WalkParams(**dict(zip(_WP_FLOATS, fparams)), first=iparams[0], last=iparams[1], center=iparams[2],
                          seed=iparams[3] & _U64, step=iparams[4] & _U64)
    G = chain_ptr.numel() - 1
    rc = _lib.lib().jamun_walk_step(_ptr(y), _ptr(v), _ptr(ybar), _ptr(p), _ptr(g), _ptr(score_in),
                                    _ptr(chain_ptr, torch.int32), G,
                                    C.byref(prm), _ptr(noise), _ptr(xhat), _ptr(score), _ptr(traj_y), _ptr(traj_xhat),
                                    _ptr(traj_score), _ptr(dev_state, torch.int64), _stream())
    _lib.check(rc, "jamun_walk_step")
    _count()


def walk_step(y, v, ybar, p, g, chain_ptr, prm: "_lib.WalkParams", noise, xhat, score, traj_y=None, traj_xhat=None,
              traj_score=None, score_in=None, dev_state=None):
    """One fused walk-jump step (jamun_walk_step) as the operator jamun_b200::baoab_step."""
    _T.baoab_step(y, v, ybar, p, g, chain_ptr, [float(getattr(prm, k)) for k in _WP_FLOATS],
                  [int(prm.first), int(prm.last), int(prm.center), _signed64(int(prm.seed)), _signed64(int(prm.step))], noise, xhat,
                  score, traj_y, traj_xhat, traj_score, score_in, dev_state)


@_op("walk_advance", mutates=("dev_state",))
def _walk_advance(dev_state: Tensor, slot_inc: int) -> None:
    rc = _lib.lib().jamun_walk_advance(_ptr(dev_state, torch.int64), int(slot_inc), _stream())
    _lib.check(rc, "jamun_walk_advance")
    _count()


def walk_advance(dev_state, slot_inc: int):
    _T.walk_advance(dev_state, int(slot_inc))


@_op("aboba_drift", mutates=("y",))
def _aboba_drift(y: Tensor, v: Tensor, half_delta: float) -> None:
    rc = _lib.lib().jamun_aboba_drift(_ptr(y), _ptr(v), float(half_delta), y.shape[0], _stream())
    _lib.check(rc, "jamun_aboba_drift")
    _count()


def aboba_drift(y, v, half_delta: float):
    _T.aboba_drift(y, v, float(half_delta))


@_op("aboba_kick", mutates=("y", "v"))
def _aboba_kick(y: Tensor, v: Tensor, score: Tensor, fparams: List[float], iparams: List[int], noise: Optional[Tensor]) -> None:
    prm = _lib.WalkParams(**dict(zip(_WP_FLOATS, fparams)), first=iparams[0], last=iparams[1], center=iparams[2],
                          seed=iparams[3] & _U64, step=iparams[4] & _U64)
    rc = _lib.lib().jamun_aboba_kick(_ptr(y), _ptr(v), _ptr(score), C.byref(prm), _ptr(noise), y.shape[0], _stream())
    _lib.check(rc, "jamun_aboba_kick")
    _count()


def aboba_kick(y, v, score, prm: "_lib.WalkParams", noise):
    _T.aboba_kick(y, v, score, [float(getattr(prm, k)) for k in _WP_FLOATS],
                  [int(prm.first), int(prm.last), int(prm.center), _signed64(int(prm.seed)), _signed64(int(prm.step))], noise)


@_op("gaussian_axpy", mutates=("out",))
def _gaussian_axpy(x: Optional[Tensor], a: float, b: float, noise: Optional[Tensor], seed: int, step: int, out: Tensor) -> None:
    n_atoms = out.shape[0]
    rc = _lib.lib().jamun_gaussian_axpy(_ptr(x), float(a), float(b), _ptr(noise), int(seed) & _U64, int(step) & _U64, n_atoms,
                                        _ptr(out), _stream())
    _lib.check(rc, "jamun_gaussian_axpy")
    _count()


def gaussian_axpy(x, a: float, b: float, noise, seed: int, step: int, out):
    _T.gaussian_axpy(x, float(a), float(b), noise, _signed64(int(seed)), _signed64(int(step)), out)
    return out


@torch.library.custom_op("jamun_b200::layout_to_soa", mutates_args=(), device_types="cuda")
def layout_to_soa(x: Tensor, s: int, v: int) -> Tensor:
    out = torch.empty_like(x)
    rc = _lib.lib().jamun_layout_to_soa(_ptr(x), s, v, x.shape[0], _ptr(out), _stream())
    _lib.check(rc, "jamun_layout_to_soa")
    _count()
    return out


@torch.library.custom_op("jamun_b200::layout_from_soa", mutates_args=(), device_types="cuda")
def layout_from_soa(x: Tensor, s: int, v: int) -> Tensor:
    out = torch.empty_like(x)
    rc = _lib.lib().jamun_layout_from_soa(_ptr(x), s, v, x.shape[0], _ptr(out), _stream())
    _lib.check(rc, "jamun_layout_from_soa")
    _count()
    return out


layout_to_soa.register_fake(lambda x, s, v: torch.empty_like(x))
layout_from_soa.register_fake(lambda x, s, v: torch.empty_like(x))


@torch.library.custom_op("jamun_b200::linear_act", mutates_args=(), device_types="cuda")
def linear_act(x: Tensor, w: Tensor, b: Tensor, act: int = 0) -> Tensor:
    out = torch.empty(x.shape[0], w.shape[0], device=x.device, dtype=torch.float32)
    rc = _lib.lib().jamun_linear_act(_ptr(x), _ptr(w), _ptr(b), x.shape[0], x.shape[1], w.shape[0], int(act), _ptr(out), _stream())
    _lib.check(rc, "jamun_linear_act")
    _count()
    return out


linear_act.register_fake(lambda x, w, b, act=0: x.new_empty(x.shape[0], w.shape[0]))


@torch.library.custom_op("jamun_b200::tensor_product", mutates_args=(), device_types="cuda")
def tensor_product(x1: Tensor, x2: Tensor, weight: Tensor, table: Tensor, d_out: int) -> Tensor:
    Z = x1.shape[0]
    out = torch.empty(Z, d_out, device=x1.device, dtype=torch.float32)
    rc = _lib.lib().jamun_tensor_product(_ptr(x1), x1.shape[1], _ptr(x2), x2.shape[1], _ptr(weight), weight.stride(0),
                                         _ptr(table, torch.int32), table.shape[0], d_out, Z, _ptr(out), _stream())
    _lib.check(rc, "jamun_tensor_product")
    _count()
    return out


# ---- training path (backward kernels); the torch.library operators built on these are in jamun_b200/autograd_ops.py --------
def _pc(t: torch.Tensor, col: int = 0) -> int:
    """Device pointer of column `col` of a row-major fp32 matrix (rows keep their leading dimension)."""
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError("jamun_b200 ops need CUDA fp32 tensors (no CPU fallback)")
    return t.data_ptr() + 4 * col


_scratch = {}


def _scratch_for(n: int, device) -> torch.Tensor:
    buf = _scratch.get(device)
    if buf is None or buf.numel() < n:
        buf = _scratch[device] = torch.empty(max(n, 1 << 16), dtype=torch.float32, device=device)
    return buf


def rowmat_mul(X, xcol: int, W, wcol: int, Y, ycol: int, a: int, b: int, trans_w: bool = False, accumulate: bool = False,
               rows: Optional[int] = None, rows_dev=None):
    rows = X.shape[0] if rows is None else rows
    rc = _lib.lib().jamun_rowmat_mul(_pc(X, xcol), X.stride(0), _pc(W, wcol), W.stride(0), int(trans_w), _pc(Y, ycol), Y.stride(0),
                                     rows, _ptr(rows_dev, torch.int32), a, b, int(accumulate), _stream())
    _lib.check(rc, "jamun_rowmat_mul")
    _count()


def rowmat_dw(X, xcol: int, dY, ycol: int, dW, wcol: int, a: int, b: int, trans_w: bool = False, accumulate: bool = False,
              rows: Optional[int] = None, rows_dev=None):
    rows = X.shape[0] if rows is None else rows
    scratch = _scratch_for(int(_lib.lib().jamun_rowmat_dw_scratch(rows, a, b)), X.device)
    rc = _lib.lib().jamun_rowmat_dw(_pc(X, xcol), X.stride(0), _pc(dY, ycol), dY.stride(0), _pc(dW, wcol), dW.stride(0), int(trans_w),
                                    rows, _ptr(rows_dev, torch.int32), a, b, int(accumulate), _ptr(scratch), _stream())
    _lib.check(rc, "jamun_rowmat_dw")
    _count(2)


def colsum(M, cols: int, out, flag=None, flag_value: int = 0, fold_s: int = 0, fold_v: int = 0, accumulate: bool = False,
           rows: Optional[int] = None, rows_dev=None, mcol: int = 0, ocol: int = 0):
    rows = M.shape[0] if rows is None else rows
    scratch = _scratch_for(int(_lib.lib().jamun_colsum_scratch(rows, cols)), M.device)
    rc = _lib.lib().jamun_colsum(_pc(M, mcol), M.stride(0), rows, _ptr(rows_dev, torch.int32), cols, _ptr(flag, torch.uint8),
                                 int(flag_value), fold_s, fold_v, out.data_ptr() + 4 * ocol, int(accumulate), _ptr(scratch), _stream())
    _lib.check(rc, "jamun_colsum")
    _count(2)


def conv_bwd_scale(dout, inv_deg, alpha0: float, alpha1: float, g):
    rc = _lib.lib().jamun_conv_bwd_scale(_ptr(dout), _ptr(inv_deg), float(alpha0), float(alpha1), dout.shape[0], _ptr(g), _stream())
    _lib.check(rc, "jamun_conv_bwd_scale")
    _count()


def stage_atb(a_ptr: int, a_comp_stride: int, ncomp: int, n_stages: int, nslots: int, rows: int, rows_pad: int, b, b_col0: int,
              b_comp_stride: int, W: int, out, mode: int, out_rows: int, slot_row0=None, slot_rows=None):
    IA = C.c_int * nslots
    r0 = IA(*slot_row0) if slot_row0 is not None else None
    rn = IA(*slot_rows) if slot_rows is not None else None
    rc = _lib.lib().jamun_stage_atb(a_ptr, int(a_comp_stride), ncomp, n_stages, nslots, rows, rows_pad, _ptr(b), b.stride(0), b_col0,
                                    b_comp_stride, W, _ptr(out), mode, out_rows, r0, rn, _stream())
    _lib.check(rc, "jamun_stage_atb")
    _count()


ATB_IMPL = None  # "tc" (tcgen05) | "simt" | "auto" (default: tc from 4096 nodes); read from JAMUN_B200_ATB at call time


def stage_atb_auto(a_ptr: int, a_comp_stride: int, ncomp: int, n_stages: int, nslots: int, rows: int, rows_pad: int, b, b_col0: int,
                   b_comp_stride: int, W: int, out, mode: int, out_rows: int, slot_row0=None, slot_rows=None):
    """dW = A^T . B over the nodes: the tensor-core kernel (jamun_stage_atb_tc) unless JAMUN_B200_ATB=simt."""
    import os

    impl = ATB_IMPL or os.environ.get("JAMUN_B200_ATB", "auto")
    if impl == "auto":  # short reductions (reference batch size 32: ~1 k nodes) do not amortise the operand split + split-K reduce
        impl = "tc" if rows >= 4096 else "simt"
    if impl == "simt":
        return stage_atb(a_ptr, a_comp_stride, ncomp, n_stages, nslots, rows, rows_pad, b, b_col0, b_comp_stride, W, out, mode, out_rows,
                         slot_row0, slot_rows)
    nslots_b = (W + 31) // 32
    per = 2 * nslots_b * rows_pad * 32
    bsplit = torch.empty(ncomp * per, dtype=torch.float32, device=b.device)
    for c in range(ncomp):
        rc = _lib.lib().jamun_pack_rows_split(_ptr(b), b.stride(0), b_col0 + c * b_comp_stride, W, rows, rows_pad, nslots_b,
                                              bsplit.data_ptr() + 4 * c * per, _stream())
        _lib.check(rc, "jamun_pack_rows_split")
    m_tiles = (n_stages + 3) // 4
    kq = ncomp * ((rows + 31) // 32)
    k_splits = max(1, min(kq, (4 * 148 + m_tiles - 1) // m_tiles, 64))  # ~4 CTAs per SM, bounded accumulation length
    partial = _scratch_for(int(_lib.lib().jamun_stage_atb_tc_scratch(n_stages, W, k_splits)), b.device)
    IA = C.c_int * nslots
    r0 = IA(*slot_row0) if slot_row0 is not None else None
    rn = IA(*slot_rows) if slot_rows is not None else None
    rc = _lib.lib().jamun_stage_atb_tc(a_ptr, int(a_comp_stride), ncomp, n_stages, nslots, rows, rows_pad, _ptr(bsplit), W, _ptr(out), mode,
                                       out_rows, r0, rn, k_splits, _ptr(partial), _stream())
    _lib.check(rc, "jamun_stage_atb_tc")
    _count(ncomp + 2)


def conv_bwd_edge(x, s_in: int, v_in: int, rowptr, col, h, rhat, dA0, dA1, dh, dxe):
    i32 = torch.int32
    rc = _lib.lib().jamun_conv_bwd_edge(_ptr(x), s_in, v_in, _ptr(rowptr, i32), _ptr(col, i32), _ptr(h), _ptr(rhat), _ptr(dA0),
                                        dA0.stride(0), _ptr(dA1), dA1.stride(1) if dA1 is not None else 0,
                                        dA1.stride(0) if dA1 is not None else 0, x.shape[0], _ptr(dh), _ptr(dxe), _stream())
    _lib.check(rc, "jamun_conv_bwd_edge")
    _count()


def conv_bwd_p2(src_rowptr, src_eid, edst, h, rhat, y, g, rows_pad: int, dh, dy_op):
    i32 = torch.int32
    N = g.shape[0]
    rc = _lib.lib().jamun_conv_bwd_p2(_ptr(src_rowptr, i32), _ptr(src_eid, i32), _ptr(edst, i32), _ptr(h), _ptr(rhat), _ptr(y),
                                      y.stride(0), _ptr(g), N, rows_pad, _ptr(dh), _ptr(dy_op), _stream())
    _lib.check(rc, "jamun_conv_bwd_p2")
    _count()


def conv_bwd_gather(src_rowptr, src_eid, dxe, extra, n_extra: int, dx):
    i32 = torch.int32
    rc = _lib.lib().jamun_conv_bwd_gather(_ptr(src_rowptr, i32), _ptr(src_eid, i32), _ptr(dxe), dx.shape[1], _ptr(extra),
                                          extra.stride(0) if extra is not None else 0, n_extra, dx.shape[0], _ptr(dx), _stream())
    _lib.check(rc, "jamun_conv_bwd_gather")
    _count()


def gate_fwd(conv, c_act: float, c_gate: float, gated):
    rc = _lib.lib().jamun_gate_fwd(_ptr(conv), float(c_act), float(c_gate), conv.shape[0], _ptr(gated), _stream())
    _lib.check(rc, "jamun_gate_fwd")
    _count()


def gate_bwd(conv, dgated, c_act: float, c_gate: float, dconv):
    rc = _lib.lib().jamun_gate_bwd(_ptr(conv), _ptr(dgated), float(c_act), float(c_gate), conv.shape[0], _ptr(dconv), _stream())
    _lib.check(rc, "jamun_gate_bwd")
    _count()


def mix_bwd(dx_new, dx_scaled, y, x_res, skip_w, s_next, dy, dx_res, prod_s, prod_w):
    rc = _lib.lib().jamun_mix_bwd(_ptr(dx_new), _ptr(dx_scaled), _ptr(y), _ptr(x_res), _ptr(skip_w), _ptr(s_next), y.shape[0],
                                  _ptr(dy), _ptr(dx_res), _ptr(prod_s), _ptr(prod_w), _stream())
    _lib.check(rc, "jamun_mix_bwd")
    _count()


def head_bwd(pre, hv, w2, dg, c_gate: float, dpre, dhv, prod_w2):
    rc = _lib.lib().jamun_head_bwd(_ptr(pre), _ptr(hv), _ptr(w2), _ptr(dg), float(c_gate), pre.shape[0], _ptr(dpre), _ptr(dhv),
                                   _ptr(prod_w2), _stream())
    _lib.check(rc, "jamun_head_bwd")
    _count()


def radial_bwd(rb, ebond, rowptr, w0r, b0eff, dh, dz):
    N, cap = rowptr.numel() - 1, ebond.numel()
    rc = _lib.lib().jamun_radial_bwd(_ptr(rb), _ptr(ebond, torch.uint8), _ptr(rowptr, torch.int32), N, cap, _ptr(w0r), _ptr(b0eff),
                                     _ptr(dh), _ptr(dz), _stream())
    _lib.check(rc, "jamun_radial_bwd")
    _count()


def embed_bwd(idx, tabs, scale, dx0, dtabs, prod):
    i32 = torch.int32
    N = dx0.shape[0]
    rc = _lib.lib().jamun_embed_bwd(_ptr(idx[0], i32), _ptr(idx[1], i32), _ptr(idx[2], i32),
                                    _ptr(idx[3], i32) if idx[3] is not None else None, *[_ptr(t) for t in tabs],
                                    *[t.shape[1] for t in tabs], *[t.shape[0] for t in tabs], _ptr(scale), _ptr(dx0), N,
                                    *[_ptr(t) for t in dtabs], _ptr(prod), _stream())
    _lib.check(rc, "jamun_embed_bwd")
    _count(5)


def noise_mlp_bwd(w1, b1, w2, b2, c_noise: float, apply_sigmoid: bool, dout, dw1, db1, dw2, db2):
    rc = _lib.lib().jamun_noise_mlp_bwd(_ptr(w1), _ptr(b1), _ptr(w2), _ptr(b2), float(c_noise), b1.numel(), int(apply_sigmoid),
                                        _ptr(dout), _ptr(dw1), _ptr(db1), _ptr(dw2), _ptr(db2), _stream())
    _lib.check(rc, "jamun_noise_mlp_bwd")
    _count()


def combine_xhat(g, ybar, chain_ptr, c_skip: float, c_mix: float, center: bool, xhat):
    G = chain_ptr.numel() - 1
    rc = _lib.lib().jamun_combine_xhat(_ptr(g), _ptr(ybar), _ptr(chain_ptr, torch.int32), G, float(c_skip), float(c_mix), int(center),
                                       _ptr(xhat), _stream())
    _lib.check(rc, "jamun_combine_xhat")
    _count()
    return xhat


def loss_fwd(xhat, x, chain_ptr, scale: float, sigma: float, lw, loss, raw, rmsd):
    G = chain_ptr.numel() - 1
    rc = _lib.lib().jamun_loss_fwd(_ptr(xhat), _ptr(x), _ptr(chain_ptr, torch.int32), G, float(scale), float(sigma), _ptr(lw),
                                   _ptr(loss), _ptr(raw), _ptr(rmsd), _stream())
    _lib.check(rc, "jamun_loss_fwd")
    _count()


def loss_bwd(xhat, x, chain_of, chain_ptr, scale: float, lw, dloss, dxhat):
    rc = _lib.lib().jamun_loss_bwd(_ptr(xhat), _ptr(x), _ptr(chain_of, torch.int32), _ptr(chain_ptr, torch.int32), xhat.shape[0],
                                   float(scale), _ptr(lw), _ptr(dloss), _ptr(dxhat), _stream())
    _lib.check(rc, "jamun_loss_bwd")
    _count()


def kabsch_align(y, x, chain_ptr, out, rot=None):
    G = chain_ptr.numel() - 1
    rc = _lib.lib().jamun_kabsch_align(_ptr(y), _ptr(x), _ptr(chain_ptr, torch.int32), G, _ptr(out), _ptr(rot), _stream())
    _lib.check(rc, "jamun_kabsch_align")
    _count()
    return out


def add_cols(out, col0: int, add, n: int):
    rc = _lib.lib().jamun_add_cols(_ptr(out), out.stride(0), col0, _ptr(add), add.stride(0), n, out.shape[0], _stream())
    _lib.check(rc, "jamun_add_cols")
    _count()


def avg_sq_dist(pos, chain_ptr, cutoff: float):
    """[G, 2] double: per chain (sum of squared pair distances below the cutoff, number of such pairs)."""
    G = chain_ptr.numel() - 1
    sums = torch.zeros(G, 2, dtype=torch.float64, device=pos.device)
    rc = _lib.lib().jamun_avg_sq_dist(_ptr(pos), _ptr(chain_ptr, torch.int32), G, float(cutoff), sums.data_ptr(), _stream())
    _lib.check(rc, "jamun_avg_sq_dist")
    _count()
    return sums


def ema_update(ema, p, decay: float):
    rc = _lib.lib().jamun_ema_update(_ptr(ema), _ptr(p), float(decay), ema.numel(), _stream())
    _lib.check(rc, "jamun_ema_update")
    _count()


tensor_product.register_fake(lambda x1, x2, weight, table, d_out: x1.new_empty(x1.shape[0], d_out))
