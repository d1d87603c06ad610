import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def make_oracle_batch(t, dtype=torch.float32):
    from oracle import jamun_oracle as O

    return O.OracleBatch(pos=t["pos"].to(dtype), batch=t["batch"], num_graphs=t["num_graphs"], edge_index=t["edge_index"],
                         atom_type_index=t["atom_type_index"], atom_code_index=t["atom_code_index"],
                         residue_code_index=t["residue_code_index"], residue_sequence_index=t["residue_sequence_index"],
                         loss_weight=t["loss_weight"].to(dtype))


@pytest.fixture(scope="session")
def models():
    """(oracle fp32, oracle fp64, product model) sharing one set of seed-0 random weights with parity re-draws."""
    import jamun_b200 as J
    from oracle import jamun_oracle as O

    torch.manual_seed(0)
    o32 = O.Denoiser()
    O.randomize_for_parity(o32)
    o64 = O.Denoiser()
    o64.load_state_dict(o32.state_dict())
    o64 = o64.double()
    prod = J.default_denoiser()
    prod.load_state_dict(o32.state_dict())
    if torch.cuda.is_available():
        prod = prod.to("cuda")
    return o32, o64, prod
