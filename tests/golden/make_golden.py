"""Generate tests/golden/walkjump_small.npz from the CPU oracle (fp32 + fp64 cross-check).

The reference cannot be imported in the build container (e3nn / torch_cluster / torch_scatter / PyG / Lightning are
absent), so these vectors pin the *oracle*, not the reference: PARITY UNPINNED.  They freeze the oracle's
behaviour so that later edits to oracle/ or to the kernels are caught, and they travel to the GPU box.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from jamun_b200 import synthetic  # noqa: E402
from oracle import jamun_oracle as O  # noqa: E402

SIZES = [22, 15, 9, 30]
SIGMA = 0.04
STEPS = 4
MCMC = dict(delta=0.04, friction=1.0, M=1.0, inverse_temperature=1.0, score_fn_clip=100.0)


def weights_digest(model) -> str:
    h = hashlib.sha256()
    for k, v in sorted(model.state_dict().items()):
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def build():
    torch.manual_seed(0)
    o = O.Denoiser()
    O.randomize_for_parity(o)
    t = synthetic.make_tensors(SIZES)
    ob = O.OracleBatch(pos=t["pos"], batch=t["batch"], num_graphs=len(SIZES), edge_index=t["edge_index"],
                       atom_type_index=t["atom_type_index"], atom_code_index=t["atom_code_index"],
                       residue_code_index=t["residue_code_index"], residue_sequence_index=t["residue_sequence_index"])
    gen = torch.Generator().manual_seed(2024)
    y0 = t["pos"] + SIGMA * torch.randn(t["pos"].shape, generator=gen)
    noise = torch.randn(STEPS, y0.shape[0], 3, generator=gen)
    it = iter(noise)
    with torch.no_grad():
        xhat = o.xhat(ob.with_pos(y0), SIGMA)
        score = o.score(ob.with_pos(y0), SIGMA)
        ybar = O.mean_center_pos(y0, ob.batch, ob.num_graphs)
        sig = torch.tensor(SIGMA)
        r_cut = o.effective_radial_cutoff(sig) / o.normalization_factors(sig, 0.332)[0]
        edges = o.add_edges(ob.with_pos(ybar), r_cut)
        wj = O.walk_jump(o, ob, y0, SIGMA, mcmc=O.baoab, v_init="gaussian", steps=STEPS, save_trajectory=True,
                         noise_fn=lambda yy: next(it), redundant_jump=True, **MCMC)
    return o, dict(sizes=np.array(SIZES), sigma=SIGMA, steps=STEPS, y0=y0.numpy(), noise=noise.numpy(), xhat=xhat.numpy(),
                   score=score.numpy(), r_cut=float(r_cut), edge_index=edges.edge_index.numpy(), bond_mask=edges.bond_mask.numpy(),
                   wj_y=wj["y"].numpy(), wj_v=wj["v"].numpy(), wj_xhat=wj["xhat"].numpy(), wj_y_traj=wj["y_traj"].numpy(),
                   wj_xhat_traj=wj["xhat_traj"].numpy(), wj_score_traj=wj["score_traj"].numpy(),
                   weights_sha256=weights_digest(o))


if __name__ == "__main__":
    _, g = build()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "walkjump_small.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes; weights", g["weights_sha256"][:16])
