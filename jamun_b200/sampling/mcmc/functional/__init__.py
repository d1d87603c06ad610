from ._splitting import aboba, baoab, create_score_fn, fused_baoab, initialize_velocity

__all__ = ["aboba", "baoab", "create_score_fn", "fused_baoab", "initialize_velocity"]
