"""CPU suite, part 1: pin the oracle.  The reference has no tests or golden vectors for this path, so
the pins are analytic known answers, symmetry properties the reference implies, and the committed fixture."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import make_oracle_batch
from jamun_b200 import synthetic
from oracle import jamun_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "walkjump_small.npz")


def test_parameter_count_and_state_dict_layout(models):
    o32 = models[0]
    assert sum(p.numel() for p in o32.parameters()) == 10_552_889
    sd = o32.state_dict()
    assert tuple(sd["g.initial_projector.gated_conv.f.f.radial_nn.3.weight"].shape) == (10304, 64)
    assert tuple(sd["g.layers.4.gated_conv.f.f.radial_nn.3.weight"].shape) == (28992, 64)
    assert tuple(sd["g.layers.0.gated_conv.skip_connection.weight"].shape) == (15424,)
    assert tuple(sd["g.initial_projector.gated_conv.skip_connection.weight"].shape) == (6720,)
    assert tuple(sd["g.output_head.0.lin.weight"].shape) == (19264,)
    assert tuple(sd["g.output_head.1.weight"].shape) == (32,)
    assert tuple(sd["g.skip_connections.2.weights.scale_predictor.2.weight"].shape) == (152, 152)


def test_normalization_constants():
    c_in, c_skip, c_out, c_noise = O.Denoiser.normalization_factors(torch.tensor(0.04), 0.332)
    assert float(c_in) == pytest.approx(1.71096, abs=1e-5)
    assert float(c_skip) == pytest.approx(0.971897, abs=1e-6)
    assert float(c_out) == pytest.approx(0.0965930, abs=1e-6)
    assert float(c_noise) == pytest.approx(-0.804719, abs=1e-6)
    r = torch.sqrt(torch.tensor(1.0) + 6 * torch.tensor(0.04) ** 2) / c_in
    assert float(r) == pytest.approx(0.587264, abs=1e-6)


def test_e3nn_constants():
    assert O.normalize2mom_const("leaky_relu") == pytest.approx(1.4162684218969974, rel=1e-12)
    assert O.normalize2mom_const("sigmoid") == pytest.approx(1.8467055342154763, rel=1e-12)
    w = O.wigner_3j(1, 1, 1)
    assert float(w[0, 1, 2]) == pytest.approx(1 / math.sqrt(6))
    assert float(w[0, 2, 1]) == pytest.approx(-1 / math.sqrt(6))
    for key in ((0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 1)):
        assert float(O.wigner_3j(*key).pow(2).sum()) == pytest.approx(1.0)


def test_radial_basis_and_sh_known_answers():
    values = torch.linspace(0.0, 0.5872643, 34)[1:-1]
    rb = O.soft_one_hot_linspace_gaussian_cutoff(values, 0.0, 0.5872643, 32)
    assert torch.allclose(torch.diagonal(rb), torch.full((32,), 1 / 1.12), atol=1e-6)
    assert float(rb[0, 1]) == pytest.approx(math.exp(-1.0) / 1.12, rel=1e-4)
    sh = O.spherical_harmonics_l01(torch.tensor([[0.0, 0.0, 2.0], [3.0, 0.0, 0.0]]))
    assert torch.allclose(sh, torch.tensor([[1.0, 0, 0, math.sqrt(3)], [1.0, math.sqrt(3), 0, 0]]))


def test_single_edge_tensor_product_by_hand():
    """FCTP on one edge reduces to hand-computable sums (SURVEY 8c iii)."""
    tp = O.FullyConnectedTP("2x0e+1x1e", "1x0e+1x1e", "1x0e+1x1e")
    assert tp.weight_numel == 2 + 2 + 1 + 1 + 1
    x = torch.tensor([[2.0, -1.0, 0.5, 1.5, -2.0]], dtype=torch.float64)
    sh = torch.tensor([[1.0, 0.3, -0.4, 1.2]], dtype=torch.float64)
    w = torch.tensor([[0.7, -0.2, 1.1, 0.4, -0.9, 0.6, 1.3]], dtype=torch.float64)
    out = tp(x, sh, w)[0]
    xs, xv, s1 = x[0, :2], x[0, 2:], sh[0, 1:]
    t1 = 0.7 * xs[0] - 0.2 * xs[1]
    t4 = 0.6 * (xv @ s1) / math.sqrt(3)
    want0 = math.sqrt(1 / 3) * (t1 + t4)
    t2 = (1.1 * xs[0] + 0.4 * xs[1]) * s1 / math.sqrt(3)
    t3 = -0.9 * xv / math.sqrt(3)
    t5 = 1.3 * torch.linalg.cross(xv, s1) / math.sqrt(6)
    want1 = math.sqrt(3 / 4) * (t2 + t3 + t5)
    assert torch.allclose(out, torch.cat([want0[None], want1]), atol=1e-12)


def test_o3_linear_by_hand():
    lin = O.O3Linear("2x0e+1x1e", "1x0e+2x1e")
    with torch.no_grad():
        lin.weight.copy_(torch.tensor([0.5, -1.5, 2.0, 3.0]))
    x = torch.tensor([[1.0, 2.0, 0.1, 0.2, 0.3]])
    out = lin(x)[0]
    assert out[0].item() == pytest.approx((0.5 * 1 - 1.5 * 2) / math.sqrt(2))
    assert torch.allclose(out[1:4], 2.0 * x[0, 2:]) and torch.allclose(out[4:], 3.0 * x[0, 2:])


def _brute_radius(pos, r, batch, cap):
    """O(N^2) pure-Python neighbour list with the torch_cluster GPU cap rule."""
    r2 = np.float32(np.float64(np.float32(r)) ** 2)
    p = pos.numpy()
    edges = []
    for i in range(len(p)):
        hits = 0
        for j in range(len(p)):
            if batch[j] != batch[i]:
                continue
            d = p[i] - p[j]
            d2 = np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2])
            if np.float32(d2) < r2:
                if j != i:
                    edges.append((j, i))
                hits += 1
                if cap is not None and hits >= cap + 1:
                    break
    return edges


@pytest.mark.parametrize("sizes,cap", [([12, 7, 1], 32), ([45], 4), ([45], None)])
def test_radius_graph_vs_bruteforce(sizes, cap):
    t = synthetic.make_tensors(sizes)
    ei = O.radius_graph(t["pos"], 0.5872643, t["batch"], cap)
    want = _brute_radius(t["pos"], 0.5872643, t["batch"].tolist(), cap)
    assert [(int(a), int(b)) for a, b in ei.T] == sorted(want, key=lambda e: (e[1], e[0]))
    if cap is not None:
        deg = torch.bincount(ei[1], minlength=sum(sizes))
        assert int(deg.max()) <= cap + 1


def test_equivariance_permutation_and_gain_zero(models):
    o32, o64, _ = models
    t = synthetic.make_tensors([14, 9])
    b = make_oracle_batch(t, torch.float64)
    gen = torch.Generator().manual_seed(0)
    y = b.pos + 0.04 * torch.randn(b.pos.shape, generator=gen, dtype=torch.float64)
    with torch.no_grad():
        x = o64.xhat(b.with_pos(y), 0.04)
        Q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen, dtype=torch.float64))
        if torch.det(Q) < 0:
            Q[:, 0] *= -1
        xr = o64.xhat(b.with_pos(y @ Q.T + 0.7), 0.04)
        assert torch.allclose(xr, x @ Q.T, atol=1e-12)
        # fresh model: output_gain = 0 => xhat = mean_center(c_skip * mean_center(y))
        torch.manual_seed(3)
        fresh = O.Denoiser().double()
        c_skip = fresh.normalization_factors(torch.tensor(0.04, dtype=torch.float64), 0.332)[1]
        yb = O.mean_center_pos(y, b.batch, 2)
        assert torch.allclose(fresh.xhat(b.with_pos(y), 0.04), O.mean_center_pos(c_skip * yb, b.batch, 2), atol=1e-14)


def test_baoab_zero_score_closed_form():
    """score == 0: y <- y + (d/2)(v + vhat), v <- vhat = e^-g v + sqrt(1-e^-2g) sqrt(u) R."""
    gen = torch.Generator().manual_seed(0)
    y0, R0, R1 = (torch.randn(5, 3, generator=gen, dtype=torch.float64) for _ in range(3))
    it = iter([R0, R1])
    y, v, y_traj, s_traj = O.baoab(y0, lambda y: torch.zeros_like(y), steps=2, v_init="gaussian", save_trajectory=True,
                                   delta=0.1, friction=0.5, M=4.0, noise_fn=lambda yy: next(it))
    u = 0.25
    v0 = math.sqrt(u) * R0
    vhat = math.exp(-0.5) * v0 + math.sqrt(1 - math.exp(-1.0)) * math.sqrt(u) * R1
    assert torch.allclose(v, vhat) and torch.allclose(y, y0 + 0.05 * (v0 + vhat))
    assert y_traj.shape == (2, 5, 3) and s_traj.shape == (2, 5, 3)


def test_kabsch_recovers_rigid_motion():
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(20, 3, generator=gen, dtype=torch.float64)
    batch = torch.tensor([0] * 12 + [1] * 8)
    Q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen, dtype=torch.float64))
    if torch.det(Q) < 0:
        Q[:, 0] *= -1
    y = x @ Q.T + torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    assert torch.allclose(O.kabsch_algorithm(y, x, batch, 2), x, atol=1e-10)


def test_golden_fixture_reproduced(models):
    """The committed fixture (tests/golden/make_golden.py) is reproduced bit-for-bit-ish by today's oracle."""
    from golden import make_golden

    g = np.load(GOLD)
    o, now = make_golden.build()
    assert str(g["weights_sha256"]) == now["weights_sha256"], "seed-0 weights differ from the fixture's"
    assert np.array_equal(g["edge_index"], now["edge_index"]) and np.array_equal(g["bond_mask"], now["bond_mask"])
    for k in ("xhat", "score", "wj_y", "wj_xhat_traj", "wj_score_traj"):
        tol = 1e-6 if "score" not in k else 1e-6 / 0.04 ** 2
        assert np.allclose(g[k], now[k], rtol=1e-5, atol=tol), k


def test_jump_is_free_for_baoab(models):
    """xhat of a saved frame equals y + sigma^2 * score of that frame -- what the fused path relies on."""
    g = np.load(GOLD)
    y, s, x = g["wj_y_traj"], g["wj_score_traj"], g["wj_xhat_traj"]
    assert np.allclose(y + 0.04 ** 2 * s, x, atol=2e-6)


# ---- pinning the [DEP-recalled] constants as far as this image allows -------------------------------------------------------
def _e3nn_so3_clebsch_gordan(l1, l2, l3):
    """Replay of e3nn 0.5.x `o3._wigner._so3_clebsch_gordan` (the generator of its wigner_3j constants) with exact SU(2)
    Clebsch-Gordan coefficients from sympy: C = real(Q1 (x) Q2 (x) conj(Q3^T) . CG_su2) / ||.||,
    Q_l = (-i)^l * (real -> complex change of basis), index order m = -l..l = e3nn's component order."""
    import numpy as np
    from sympy.physics.quantum.cg import CG

    def q(l):
        m_ = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
        for m in range(-l, 0):
            m_[l + m, l + abs(m)] = 1 / 2 ** 0.5
            m_[l + m, l - abs(m)] = -1j / 2 ** 0.5
        m_[l, l] = 1
        for m in range(1, l + 1):
            m_[l + m, l + abs(m)] = (-1) ** m / 2 ** 0.5
            m_[l + m, l - abs(m)] = 1j * (-1) ** m / 2 ** 0.5
        return (-1j) ** l * m_

    su2 = np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1))
    for m1 in range(-l1, l1 + 1):
        for m2 in range(-l2, l2 + 1):
            if abs(m1 + m2) <= l3:
                su2[l1 + m1, l2 + m2, l3 + m1 + m2] = float(CG(l1, m1, l2, m2, l3, m1 + m2).doit())
    c = np.einsum("ij,kl,mn,ikn->jlm", q(l1), q(l2), np.conj(q(l3).T), su2.astype(np.complex128))
    assert np.abs(c.imag).max() < 1e-12
    return c.real / np.linalg.norm(c.real)


def test_wigner_3j_matches_sympy_replay_of_e3nn_recipe():
    """The single convention risk for released checkpoints: sign / index order of wigner_3j(1,1,1)."""
    pytest.importorskip("sympy")
    for key in ((0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 1)):
        ref = torch.from_numpy(_e3nn_so3_clebsch_gordan(*key))
        assert torch.allclose(O.wigner_3j(*key), ref, atol=1e-14), key


def test_normalize2mom_recipe_is_recomputed_not_hard_coded():
    """c = E[f(z)^2]^(-1/2) over z = randn(1e6, seed-0 CPU generator, fp64) (e3nn.math.normalize2mom), recomputed here."""
    z = torch.randn(1_000_000, generator=torch.Generator("cpu").manual_seed(0), dtype=torch.float64)
    for name, f in (("leaky_relu", lambda t: torch.nn.functional.leaky_relu(t, 0.01)), ("sigmoid", torch.sigmoid)):
        c = f(z).pow(2).mean().pow(-0.5).item()
        assert O.normalize2mom_const(name) == pytest.approx(c, rel=1e-12)


# ---- dormant until e3nn is importable (e.g. a driver-provided baseline/_ref): the oracle's layers against the real ones -----
def _e3nn():
    import sys

    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    return pytest.importorskip("e3nn")


def test_oracle_tensor_product_vs_e3nn():
    e3nn = _e3nn()
    from e3nn import o3

    for in1 in ("120x0e+32x1e", "8x0e+8x0e+32x0e+8x0e"):
        tp_ref = o3.FullyConnectedTensorProduct(in1, "1x0e+1x1e", "152x0e+32x1e", shared_weights=False, internal_weights=False)
        tp = O.FullyConnectedTP(in1, "1x0e+1x1e", "152x0e+32x1e")
        assert tp.weight_numel == tp_ref.weight_numel
        gen = torch.Generator().manual_seed(0)
        x = torch.randn(5, o3.Irreps(in1).dim, generator=gen)
        sh = torch.randn(5, 4, generator=gen)
        w = torch.randn(5, tp.weight_numel, generator=gen)
        assert torch.allclose(tp(x, sh, w), tp_ref(x, sh, w), rtol=1e-5, atol=1e-5)


def test_oracle_linear_gate_radial_sh_vs_e3nn():
    e3nn = _e3nn()
    from e3nn import nn as enn
    from e3nn import o3
    from e3nn.math import soft_one_hot_linspace

    gen = torch.Generator().manual_seed(1)
    lin_ref = o3.Linear("120x0e+32x1e", "152x0e+32x1e")
    lin = O.O3Linear("120x0e+32x1e", "152x0e+32x1e")
    with torch.no_grad():
        lin.weight.copy_(lin_ref.weight)
    x = torch.randn(7, 216, generator=gen)
    assert torch.allclose(lin(x), lin_ref(x), rtol=1e-5, atol=1e-5)
    gate_ref = enn.Gate("120x0e", [torch.nn.LeakyReLU(0.01)], "32x0e", [torch.sigmoid], "32x1e")
    gate = O.Gate("120x0e+32x1e")
    xg = torch.randn(7, 248, generator=gen)
    assert torch.allclose(gate(xg), gate_ref(xg), rtol=1e-5, atol=1e-6)
    d = torch.rand(9, generator=gen) * 0.8
    rb_ref = soft_one_hot_linspace(d, 0.0, 0.5872643, 32, basis="gaussian", cutoff=True)
    assert torch.allclose(O.soft_one_hot_linspace_gaussian_cutoff(d, 0.0, 0.5872643, 32), rb_ref, rtol=1e-5, atol=1e-6)
    v = torch.randn(9, 3, generator=gen)
    sh_ref = o3.spherical_harmonics(o3.Irreps("1x0e+1x1e"), v, normalize=True, normalization="component")
    assert torch.allclose(O.spherical_harmonics_l01(v), sh_ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(O.wigner_3j(1, 1, 1), o3.wigner_3j(1, 1, 1).double(), atol=1e-7)
