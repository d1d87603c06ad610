"""ctypes binding of the jamun_b200 C ABI (include/jamun_b200.h).

The product has no CPU fallback: if the CUDA library is absent and cannot be built, importing any
compute op raises.  The library is built in-tree by jamun_b200/csrc/build.py (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "csrc" / "libjamun_b200.so"
_lib = None

c_f = C.c_void_p  # device pointers travel as integers
I, F, ULL = C.c_int, C.c_float, C.c_ulonglong


class WalkParams(C.Structure):
    _fields_ = [("c_in", F), ("c_skip", F), ("c_out", F), ("sigma2", F),
                ("delta", F), ("u", F), ("a", F), ("z_sqrt_u", F), ("beta", F), ("clip", F),
                ("first", I), ("last", I), ("center", I), ("seed", ULL), ("step", ULL)]


class GemmEpilogue(C.Structure):
    """jamun_gemm_epilogue (include/jamun_b200.h): fused ConvBlock epilogues of jamun_gemm_f16x3_fused."""
    _fields_ = [("mode", I), ("op_s", C.c_void_p), ("op_v", C.c_void_p), ("op_v_comp_stride", C.c_longlong), ("op_rows_pad", I),
                ("c_act", F), ("c_gate", F), ("x_res", C.c_void_p), ("skip_w", C.c_void_p), ("s_next", C.c_void_p),
                ("x_new", C.c_void_p), ("x_scaled", C.c_void_p)]


_PROTOS = {
    "jamun_abi_version": ([], I),
    "jamun_last_error": ([], C.c_char_p),
    "jamun_noise_mlp": ([c_f, c_f, c_f, c_f, F, I, I, c_f, c_f], I),
    "jamun_atom_embed": ([c_f] * 8 + [I] * 4 + [c_f, I, c_f, c_f], I),
    "jamun_center_scale": ([c_f, c_f, I, I, F, c_f, c_f, c_f], I),
    "jamun_radius_csr": ([c_f, c_f, c_f, I, F, I, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f], I),
    "jamun_radius_csr_cells": ([c_f, c_f, I, I, I, F, F, I, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f], I),
    "jamun_edge_geom": ([c_f, c_f, c_f, c_f, I, I, c_f, F, c_f, c_f, c_f], I),
    "jamun_edge_radial_hidden": ([c_f, c_f, c_f, I, I, c_f, c_f, c_f, c_f], I),
    "jamun_edge_radial_hidden_all": ([c_f, c_f, c_f, I, I, c_f, c_f, I, c_f, c_f], I),
    "jamun_radial_pack_frag": ([c_f, I, c_f, c_f], I),
    "jamun_edge_radial_hidden_mma": ([c_f, c_f, c_f, I, I, c_f, c_f, I, c_f, c_f], I),
    "jamun_conv_fwd": ([c_f, I, I, c_f, c_f, c_f, c_f, c_f, c_f, F, F, I, c_f, c_f], I),
    "jamun_conv_build_a": ([c_f, I, I, c_f, c_f, c_f, c_f, c_f, I, I, I, I, c_f, c_f, C.c_longlong, c_f, I, F, c_f, c_f], I),
    "jamun_conv_build_tc": ([c_f, I, I, c_f, c_f, c_f, c_f, I, I, I, c_f, c_f, C.c_longlong, c_f, c_f], I),
    "jamun_conv_build_tc_tiled": ([c_f, I, I, c_f, c_f, c_f, c_f, I, I, I, c_f, c_f, C.c_longlong, c_f, c_f], I),
    "jamun_conv_p2": ([c_f, c_f, c_f, c_f, c_f, c_f, I, c_f, c_f, I, F, c_f, c_f], I),
    "jamun_csr_by_source": ([c_f, c_f, I, I, c_f, c_f, c_f, c_f], I),
    "jamun_pack_b": ([c_f, I, c_f, I, I, I, I, I, I, I, I, c_f, c_f], I),
    "jamun_tensor_product": ([c_f, I, c_f, I, c_f, C.c_longlong, c_f, I, I, I, c_f, c_f], I),
    "jamun_rowmat_mul": ([c_f, I, c_f, I, I, c_f, I, I, c_f, I, I, I, c_f], I),
    "jamun_rowmat_dw_scratch": ([I, I, I], C.c_longlong),
    "jamun_rowmat_dw": ([c_f, I, c_f, I, c_f, I, I, I, c_f, I, I, I, c_f, c_f], I),
    "jamun_colsum_scratch": ([I, I], C.c_longlong),
    "jamun_colsum": ([c_f, I, I, c_f, I, c_f, I, I, I, c_f, I, c_f, c_f], I),
    "jamun_conv_bwd_scale": ([c_f, c_f, F, F, I, c_f, c_f], I),
    "jamun_stage_atb": ([c_f, C.c_longlong, I, I, I, I, I, c_f, I, I, I, I, c_f, I, I, C.POINTER(I), C.POINTER(I), c_f], I),
    "jamun_pack_rows_split": ([c_f, I, I, I, I, I, I, c_f, c_f], I),
    "jamun_stage_atb_tc_scratch": ([I, I, I], C.c_longlong),
    "jamun_stage_atb_tc": ([c_f, C.c_longlong, I, I, I, I, I, c_f, I, c_f, I, I, C.POINTER(I), C.POINTER(I), I, c_f, c_f], I),
    "jamun_conv_bwd_edge": ([c_f, I, I, c_f, c_f, c_f, c_f, c_f, I, c_f, I, C.c_longlong, I, c_f, c_f, c_f], I),
    "jamun_conv_bwd_p2": ([c_f, c_f, c_f, c_f, c_f, c_f, I, c_f, I, I, c_f, c_f, c_f], I),
    "jamun_conv_bwd_gather": ([c_f, c_f, c_f, I, c_f, I, I, I, c_f, c_f], I),
    "jamun_gate_fwd": ([c_f, F, F, I, c_f, c_f], I),
    "jamun_gate_bwd": ([c_f, c_f, F, F, I, c_f, c_f], I),
    "jamun_mix_bwd": ([c_f, c_f, c_f, c_f, c_f, c_f, I, c_f, c_f, c_f, c_f, c_f], I),
    "jamun_head_bwd": ([c_f, c_f, c_f, c_f, F, I, c_f, c_f, c_f, c_f], I),
    "jamun_radial_bwd": ([c_f, c_f, c_f, I, I, c_f, c_f, c_f, c_f, c_f], I),
    "jamun_embed_bwd": ([c_f] * 8 + [I] * 8 + [c_f, c_f, I, c_f, c_f, c_f, c_f, c_f, c_f], I),
    "jamun_noise_mlp_bwd": ([c_f, c_f, c_f, c_f, F, I, I, c_f, c_f, c_f, c_f, c_f, c_f], I),
    "jamun_combine_xhat": ([c_f, c_f, c_f, I, F, F, I, c_f, c_f], I),
    "jamun_loss_fwd": ([c_f, c_f, c_f, I, F, F, c_f, c_f, c_f, c_f, c_f], I),
    "jamun_loss_bwd": ([c_f, c_f, c_f, c_f, I, F, c_f, c_f, c_f, c_f], I),
    "jamun_kabsch_align": ([c_f, c_f, c_f, I, c_f, c_f, c_f], I),
    "jamun_avg_sq_dist": ([c_f, c_f, I, F, c_f, c_f], I),
    "jamun_ema_update": ([c_f, c_f, F, C.c_longlong, c_f], I),
    "jamun_add_cols": ([c_f, I, I, c_f, I, I, I, c_f], I),
    "jamun_pack_rows": ([c_f, I, I, I, I, I, c_f, c_f], I),
    "jamun_gemm_tf32x3": ([I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(I), C.POINTER(I), C.POINTER(I), C.POINTER(I),
                           C.POINTER(F), C.POINTER(C.c_void_p), C.POINTER(I), I, C.c_longlong, I, I, c_f, c_f, I, c_f], I),
    "jamun_gemm_tf32x3_splitk": ([I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(I), C.POINTER(I), C.POINTER(I), C.POINTER(I),
                                  C.POINTER(F), I, I, c_f, c_f, I, I, c_f, c_f], I),
    "jamun_gemm_f16x3": ([I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(I), C.POINTER(I), C.POINTER(I), C.POINTER(I),
                          C.POINTER(F), C.POINTER(C.c_void_p), C.POINTER(I), C.POINTER(F), I, C.c_longlong, I, I, c_f, c_f, I, I, c_f, c_f, I, c_f], I),
    "jamun_gemm_f16x3_fused": ([I, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(I), C.POINTER(I), C.POINTER(I), C.POINTER(I),
                                C.POINTER(F), C.POINTER(C.c_void_p), C.POINTER(I), C.POINTER(F), I, I, c_f, c_f, I,
                                C.POINTER(GemmEpilogue), c_f], I),
    "jamun_pack_b_f16": ([c_f, I, c_f, I, I, I, I, I, I, I, I, F, c_f, c_f], I),
    "jamun_block_tail": ([c_f, c_f, c_f, I, I, c_f, c_f, c_f, c_f, c_f, c_f, c_f, F, F, I, c_f, c_f, c_f], I),
    "jamun_tail_pack": ([c_f, c_f, c_f, I, I, F, F, I, I, c_f, c_f, C.c_longlong, c_f, c_f, c_f, F, I, c_f], I),
    "jamun_tail_mix": ([c_f, c_f, c_f, c_f, I, c_f, c_f, c_f, I, c_f], I),
    "jamun_head": ([c_f, c_f, c_f, c_f, F, I, c_f, c_f], I),
    "jamun_walk_step": ([c_f, c_f, c_f, c_f, c_f, c_f, c_f, I, C.POINTER(WalkParams), c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f], I),
    "jamun_walk_advance": ([c_f, I, c_f], I),
    "jamun_aboba_drift": ([c_f, c_f, F, I, c_f], I),
    "jamun_aboba_kick": ([c_f, c_f, c_f, C.POINTER(WalkParams), c_f, I, c_f], I),
    "jamun_gaussian_axpy": ([c_f, F, F, c_f, ULL, ULL, I, c_f, c_f], I),
    "jamun_linear_act": ([c_f, c_f, c_f, I, I, I, I, c_f, c_f], I),
    "jamun_layout_to_soa": ([c_f, I, I, I, c_f, c_f], I),
    "jamun_layout_from_soa": ([c_f, I, I, I, c_f, c_f], I),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)


def lib() -> C.CDLL:
    """Load (building first if the .so is missing and nvcc is available).  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        from .csrc import build as _build  # noqa: WPS433

        try:
            _build.build()
        except Exception as exc:  # pragma: no cover - environment dependent
            raise RuntimeError(
                f"jamun_b200: CUDA library {LIB_PATH} is missing and could not be built ({exc}). "
                "Run `python -c 'import __graft_entry__ as g; g.build()'`.  There is no CPU fallback."
            ) from exc
    handle = C.CDLL(str(LIB_PATH))
    for name, (argtypes, restype) in _PROTOS.items():
        fn = getattr(handle, name)  # AttributeError if a declared symbol is not exported
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().jamun_last_error().decode()
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
