"""Equivariant layers of the walk-jump path (mirror of jamun.e3tools.nn for e3conv.yaml's factories)."""
from ._conv import Conv, ConvBlock
from ._gate import Gate, Gated
from ._interaction import LinearSelfInteraction
from ._linear import Linear
from ._mlp import EquivariantMLP, EquivariantMLPBlock, ScalarMLP
from ._tensor_product import FullyConnectedTensorProduct

__all__ = ["Conv", "ConvBlock", "Gate", "Gated", "LinearSelfInteraction", "Linear", "EquivariantMLP",
           "EquivariantMLPBlock", "ScalarMLP", "FullyConnectedTensorProduct"]
