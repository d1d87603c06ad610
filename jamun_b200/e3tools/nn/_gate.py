"""Gate / Gated (mirror of /root/reference/src/jamun/e3tools/nn/_gate.py:10-110; e3nn.nn.Gate semantics)."""
from __future__ import annotations

import torch

from ...irreps import Irrep, Irreps

# e3nn.math.normalize2mom constants for the default activations: (E_z f(z)^2)^-1/2 over
# z = torch.randn(1_000_000, generator=Generator('cpu').manual_seed(0), dtype=float64).
# tests/test_host.py recomputes them.
C_LEAKY_RELU = 1.4162684218969974
C_SIGMOID = 1.8467055342154763


def normalize2mom_const(fn) -> float:
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    c = fn(z).pow(2).mean().pow(-0.5).item()
    return 1.0 if abs(c - 1.0) < 1e-4 else c


class Gate(torch.nn.Module):
    def __init__(self, irreps_out, act=None, act_gates=None):
        super().__init__()
        if act is not None or act_gates is not None:
            raise NotImplementedError("only the default LeakyReLU / Sigmoid gate is built into the kernels")
        self.irreps_out = Irreps(irreps_out)
        scalars = Irreps([(m, ir) for m, ir in self.irreps_out if ir.l == 0])
        gated = Irreps([(m, ir) for m, ir in self.irreps_out if ir.l > 0])
        gates = Irreps([(m, Irrep(0, 1)) for m, _ in gated])
        self.irreps_scalars, self.irreps_gates, self.irreps_gated = scalars, gates, gated
        self.irreps_in = (scalars + gates + gated).simplify()
        self.c_act, self.c_gate = C_LEAKY_RELU, C_SIGMOID

    def forward(self, x):
        """Module-level compatibility forward: [scalars | gates | gated] -> [act(scalars) | gated * act(gates)] (elementwise glue)."""
        ns, ng = self.irreps_scalars.dim, self.irreps_gates.dim
        s = torch.nn.functional.leaky_relu(x[:, :ns], 0.01) * self.c_act
        g = torch.sigmoid(x[:, ns:ns + ng]) * self.c_gate
        v = x[:, ns + ng:]
        outs, og, ov = [s], 0, 0
        for m, ir in self.irreps_gated:
            outs.append((v[:, ov:ov + m * ir.dim].reshape(-1, m, ir.dim) * g[:, og:og + m, None]).reshape(x.shape[0], -1))
            og += m
            ov += m * ir.dim
        return torch.cat(outs, dim=1)


class Gated(torch.nn.Module):
    def __init__(self, layer, irreps_in, irreps_out, act=None, act_gates=None):
        super().__init__()
        self.irreps_in, self.irreps_out = Irreps(irreps_in), Irreps(irreps_out)
        self.gate = Gate(self.irreps_out, act=act, act_gates=act_gates)
        self.f = layer(irreps_in=self.irreps_in, irreps_out=self.gate.irreps_in)
        self.irreps_sh = self.f.irreps_sh

    def forward(self, *args, **kwargs):
        return self.gate(self.f(*args, **kwargs))
