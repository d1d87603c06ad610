"""f-rows: PDB topologies in, SaveTrajectory layout out, the jamun_sample command line, average squared distance, EMA."""
import os

import numpy as np
import pytest
import torch

# ACE-ALA-GLY-NME, heavy atoms + a few hydrogens + one water (hydrogens / water must be dropped)
_ATOMS = [("CH3", "ACE", 1, "C"), ("C", "ACE", 1, "C"), ("O", "ACE", 1, "O"), ("H1", "ACE", 1, "H"),
          ("N", "ALA", 2, "N"), ("H", "ALA", 2, "H"), ("CA", "ALA", 2, "C"), ("CB", "ALA", 2, "C"), ("C", "ALA", 2, "C"),
          ("O", "ALA", 2, "O"), ("N", "GLY", 3, "N"), ("CA", "GLY", 3, "C"), ("C", "GLY", 3, "C"), ("O", "GLY", 3, "O"),
          ("N", "NME", 4, "N"), ("CH3", "NME", 4, "C"), ("HH31", "NME", 4, "H")]


def _pdb_text(seed=0):
    rng = np.random.default_rng(seed)
    lines, k = [], 1
    for name, res, seq, el in _ATOMS:
        x, y, z = rng.normal(size=3) * 3.0
        nm = name if len(name) == 4 else f" {name:<3s}"
        lines.append(f"ATOM  {k:5d} {nm} {res:>3s} A{seq:4d}    {x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00          {el:>2s}")
        k += 1
    lines.append(f"HETATM{k:5d}  O   HOH A   5       1.000   2.000   3.000  1.00  0.00           O")
    return "\n".join(lines) + "\nEND\n"


def test_pdb_reader_matches_preprocess_topology_contract():
    from jamun_b200 import pdb
    from jamun_b200.utils import ResidueMetadata, encode_atom_code, encode_atom_type, encode_residue

    g, top = pdb.graph_from_pdb(_pdb_text(), label="capped_AG")
    heavy = [a for a in _ATOMS if a[3] != "H"]
    assert top.n_atoms == len(heavy) == 14 and g.pos.shape == (14, 3)
    assert g.atom_type_index.dtype == torch.int32 and g.edge_index.dtype == torch.long
    assert g.atom_type_index.tolist() == [ResidueMetadata.ATOM_TYPES.index(a[3]) for a in heavy]
    assert g.residue_code_index.tolist() == [ResidueMetadata.RESIDUE_CODES.index(a[1]) for a in heavy]
    assert g.atom_code_index.tolist() == [encode_atom_code(a[0]) for a in heavy]
    assert encode_atom_code("CH3") == len(ResidueMetadata.ATOM_CODES) and encode_atom_type("Zn") == 5 and encode_residue("XYZ") == 22
    assert g.residue_sequence_index.tolist() == [0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3] and g.num_residues == 4
    names = [a[0] + str(a[2]) for a in heavy]
    bonds = {frozenset((names[i], names[j])) for i, j in g.edge_index.T.tolist()}
    want = {("CH31", "C1"), ("C1", "O1"), ("C1", "N2"), ("N2", "CA2"), ("CA2", "CB2"), ("CA2", "C2"), ("C2", "O2"), ("C2", "N3"),
            ("N3", "CA3"), ("CA3", "C3"), ("C3", "O3"), ("C3", "N4"), ("N4", "CH34")}
    assert bonds == {frozenset(b) for b in want}
    assert g.edge_index.shape == (2, 13)  # single direction, as data/_mdtraj.py:73
    assert float(g.loss_weight) == 1.0 and g.dataset_label == "capped_AG"


def test_pdb_dcd_npy_writers_round_trip(tmp_path):
    from jamun_b200 import pdb

    g, top = pdb.graph_from_pdb(_pdb_text(1))
    frames = g.pos.numpy()[None] + 0.01 * np.random.default_rng(2).normal(size=(5, top.n_atoms, 3)).astype(np.float32)
    pdb.write_pdb(str(tmp_path / "t.pdb"), top, frames)
    top2, xyz2 = pdb.read_pdb(str(tmp_path / "t.pdb"))
    assert [a.name for a in top2.atoms] == [a.name for a in top.atoms] and sorted(top2.bonds) == sorted(top.bonds)
    assert np.allclose(xyz2, frames[0], atol=1e-4)  # PDB keeps 1e-3 Angstrom
    pdb.write_dcd(str(tmp_path / "t.dcd"), frames)
    assert np.allclose(pdb.read_dcd(str(tmp_path / "t.dcd")), frames, atol=1e-6)


def test_save_trajectory_layout(tmp_path):
    from jamun_b200 import data, pdb
    from jamun_b200.callbacks import SaveTrajectoryCallback
    from jamun_b200.metrics import SaveTrajectory

    g, top = pdb.graph_from_pdb(_pdb_text(3), label="pep")
    metric = SaveTrajectory("pep", top, output_root=str(tmp_path), init_positions_nm=g.pos.numpy())
    cb = SaveTrajectoryCallback({"pep": metric})
    cb.on_sample_start()
    for batch in range(2):
        samples = []
        for chain in range(2):
            s = data.DataWithResidueInformation(xhat_traj=torch.randn(top.n_atoms, 4, 3), dataset_label="pep")
            samples.append(s)
        cb.on_after_sample_batch(samples)
    cb.on_sample_end()
    root = tmp_path / "sampler" / "pep"
    assert (root / "topology.pdb").exists()
    for ext in ("npy", "pdb", "dcd"):
        assert sorted(os.listdir(root / "predicted_samples" / ext)) == sorted([f"{k}.{ext}" for k in range(4)] + [f"joined.{ext}"])
    assert np.load(root / "predicted_samples" / "npy" / "2.npy").shape == (top.n_atoms, 4, 3)
    joined = np.load(root / "predicted_samples" / "npy" / "joined.npy")
    assert joined.shape == (top.n_atoms, 16, 3)
    assert np.array_equal(joined[:, 8:12], np.load(root / "predicted_samples" / "npy" / "2.npy"))
    assert pdb.read_dcd(str(root / "predicted_samples" / "dcd" / "joined.dcd")).shape == (16, top.n_atoms, 3)
    with pytest.raises(ValueError):
        metric.update(data.DataWithResidueInformation(xhat_traj=torch.randn(3, 4, 3)))


def test_cli_parses_flags_and_hydra_style_overrides(tmp_path):
    from jamun_b200.cmdline import sample as cli

    p = tmp_path / "a.pdb"
    p.write_text(_pdb_text())
    a = cli.parse_args(["--pdb", str(p), "--random-init", "num_batches=3", "continue_chain=false", "+batch_sampler.mcmc.delta=0.02",
                        "num_sampling_steps_per_batch=50", "seed=7"])
    assert (a.num_batches, a.continue_chain, a.delta, a.steps, a.seed) == (3, False, 0.02, 50, 7)
    with pytest.raises(SystemExit):
        cli.parse_args(["--random-init"])  # no initial structure
    with pytest.raises(SystemExit):
        cli.parse_args(["--pdb", str(p)])  # neither checkpoint nor --random-init
    with pytest.raises(SystemExit):
        cli.parse_args(["--pdb", str(p), "--random-init", "nonsense=1"])


@pytest.mark.gpu
def test_cli_end_to_end_random_init_with_finetuning(tmp_path):
    from jamun_b200.cmdline import sample as cli

    p = tmp_path / "pep.pdb"
    # realistic geometry: a synthetic chain's coordinates (nm -> Angstrom) under the capped-AG topology
    from jamun_b200 import synthetic

    pos = synthetic.make_chain(14, 0)["pos"] * 10.0
    rows = [a for a in _ATOMS if a[3] != "H"]
    lines = []
    for k, ((name, res, seq, el), xyz) in enumerate(zip(rows, pos), start=1):
        nm = name if len(name) == 4 else f" {name:<3s}"
        lines.append(f"ATOM  {k:5d} {nm} {res:>3s} A{seq:4d}    {xyz[0]:8.3f}{xyz[1]:8.3f}{xyz[2]:8.3f}  1.00  0.00          {el:>2s}")
    p.write_text("\n".join(lines) + "\nEND\n")
    metrics = cli.run(cli.parse_args(["--pdb", str(p), "--random-init", "--steps", "6", "--num-batches", "2", "--repeat-init-samples",
                                      "3", "--save-every-n-steps", "2", "--output-dir", str(tmp_path), "--finetune-steps", "2"]))
    root = tmp_path / "sampler" / "pep" / "predicted_samples" / "npy"
    arr = np.load(root / "0.npy")
    assert arr.shape == (14, 3, 3) and np.isfinite(arr).all()  # frames 0, 2, 4
    assert np.load(root / "joined.npy").shape == (14, 18, 3)   # 2 batches x 3 chains x 3 frames
    assert len(metrics["pep"].samples) == 6


@pytest.mark.gpu
def test_average_squared_distance_and_ema_kernels():
    from jamun_b200 import synthetic
    from jamun_b200.callbacks import EMA
    from jamun_b200.utils import compute_average_squared_distance

    t = synthetic.make_tensors([22, 15, 40])
    pos = t["pos"]
    for cutoff in (None, 0.6):
        want = []
        for c in range(3):
            x = pos[t["batch"] == c].double().numpy()
            d = np.linalg.norm(x[:, None] - x[None], axis=-1)
            m = np.tri(len(x), len(x), k=-1, dtype=bool)
            if cutoff is not None:
                m &= d < cutoff
            want.append((d[m] ** 2).mean())
        got1 = compute_average_squared_distance(pos[t["batch"] == 0].cuda(), cutoff)
        assert got1 == pytest.approx(want[0], rel=1e-5)
        got = compute_average_squared_distance(pos.cuda(), cutoff, chain_ptr=t["ptr"].to("cuda", torch.int32))
        assert got == pytest.approx(float(np.mean(want)), rel=1e-5)
    lin = torch.nn.Linear(64, 64).cuda()
    ema = EMA(decay=0.9)
    ema.on_fit_start(lin)
    ref = [p.detach().clone() for p in lin.parameters()]
    for step in range(3):
        with torch.no_grad():
            for p in lin.parameters():
                p.add_(0.1 * (step + 1))
        ema.on_train_batch_end()
        ref = [0.9 * r + 0.1 * p.detach() for r, p in zip(ref, lin.parameters())]
    for e, r in zip(ema._ema, ref):
        assert torch.allclose(e, r, rtol=1e-6, atol=1e-7)
    before = [p.detach().clone() for p in lin.parameters()]
    with ema.swapped():
        assert all(torch.allclose(p, r, rtol=1e-6, atol=1e-7) for p, r in zip(lin.parameters(), ref))
    assert all(torch.equal(p, b) for p, b in zip(lin.parameters(), before))
