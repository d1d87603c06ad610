"""Minimal irreps bookkeeping replacing e3nn.o3.Irreps on this path (mul, l, parity, dims, slices).

Follows the subset of e3nn 0.5.4's Irreps API that the reference touches
(/root/reference/src/jamun/model/arch/e3conv.py:35-37, e3tools/nn/_gate.py:49-51).
"""
from __future__ import annotations

from typing import Iterable, List, Tuple, Union


class Irrep:
    __slots__ = ("l", "p")

    def __init__(self, l: int, p: int):
        self.l, self.p = int(l), int(p)

    @property
    def dim(self) -> int:
        return 2 * self.l + 1

    def __eq__(self, other):
        return isinstance(other, Irrep) and (self.l, self.p) == (other.l, other.p)

    def __hash__(self):
        return hash((self.l, self.p))

    def __repr__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    def __mul__(self, other: "Irrep") -> List["Irrep"]:
        return [Irrep(l, self.p * other.p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]


class Irreps:
    def __init__(self, spec: Union[str, "Irreps", Iterable] = ""):
        items: List[Tuple[int, Irrep]] = []
        if isinstance(spec, Irreps):
            items = list(spec._items)
        elif isinstance(spec, str):
            for tok in spec.split("+"):
                tok = tok.strip()
                if not tok:
                    continue
                mul, ir = tok.split("x") if "x" in tok else ("1", tok)
                items.append((int(mul), Irrep(int(ir[:-1]), {"e": 1, "o": -1}[ir[-1]])))
        else:
            for it in spec:
                mul, ir = it[0], it[1]
                if not isinstance(ir, Irrep):
                    ir = Irrep(*ir) if isinstance(ir, (tuple, list)) else Irreps(f"1x{ir}")._items[0][1]
                items.append((int(mul), ir))
        self._items = items

    def __iter__(self):
        return iter(self._items)

    def __len__(self):
        return len(self._items)

    def __getitem__(self, i):
        return self._items[i]

    def __add__(self, other):
        return Irreps(self._items + Irreps(other)._items)

    def __eq__(self, other):
        try:
            return self._items == Irreps(other)._items
        except Exception:
            return False

    def __repr__(self):
        return "+".join(f"{m}x{ir}" for m, ir in self._items)

    @property
    def dim(self) -> int:
        return sum(m * ir.dim for m, ir in self._items)

    @property
    def num_irreps(self) -> int:
        return sum(m for m, _ in self._items)

    @property
    def lmax(self) -> int:
        return max((ir.l for _, ir in self._items), default=0)

    def slices(self) -> List[slice]:
        out, o = [], 0
        for m, ir in self._items:
            out.append(slice(o, o + m * ir.dim))
            o += m * ir.dim
        return out

    def simplify(self) -> "Irreps":
        out: List[Tuple[int, Irrep]] = []
        for m, ir in self._items:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + m, ir)
            else:
                out.append((m, ir))
        return Irreps(out)

    def scalars_vectors(self) -> Tuple[int, int]:
        """(#0e, #1e) multiplicities; raises if anything else is present or scalars follow vectors."""
        s = v = 0
        seen_v = False
        for m, ir in self._items:
            if ir == Irrep(0, 1) and not seen_v:
                s += m
            elif ir == Irrep(1, 1):
                v += m
                seen_v = True
            else:
                raise NotImplementedError(f"irreps {self} are outside the B200 kernels' scope (Sx0e + Vx1e)")
        return s, v
