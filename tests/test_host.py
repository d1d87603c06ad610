"""CPU suite, part 2: host logic of the product -- weight re-layout, API surface, error conventions, the C-ABI
library's exported symbols -- with no compute calls (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import kernel_model as KM
from conftest import ROOT, make_oracle_batch
from jamun_b200 import synthetic
from oracle import jamun_oracle as O


def test_library_exports_every_declared_symbol():
    from jamun_b200 import _lib

    header = open(os.path.join(ROOT, "include", "jamun_b200.h")).read()
    declared = set(re.findall(r"\b(jamun_[a-z0-9_]+)\s*\(", header))
    declared -= {"jamun_stream_t"}
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(handle, name), name
    assert _lib.lib().jamun_abi_version() == 1


def test_abi_argument_checks_return_codes_and_messages():
    """Entry points validate their arguments before touching the device: JAMUN_EINVAL (1) + text in jamun_last_error()."""
    from jamun_b200 import _lib

    lib = _lib.lib()
    fake = 0x1000  # never dereferenced: the checks below fail before any launch
    rc = lib.jamun_conv_build_tc(fake, 7, 3, fake, fake, fake, fake, 0, 4, 128, fake, fake, 0, None, None)
    assert rc != 0 and b"unsupported input irreps 7x0e+3x1e" in lib.jamun_last_error()
    rc = lib.jamun_conv_build_tc(None, 120, 32, fake, fake, fake, fake, 0, 4, 128, fake, fake, 0, None, None)
    assert rc != 0 and b"null argument" in lib.jamun_last_error()
    rc = lib.jamun_conv_build_tc(fake, 120, 32, fake, fake, fake, fake, 0, 256, 128, fake, fake, 0, None, None)
    assert rc != 0 and b"nrows exceeds rows_pad" in lib.jamun_last_error()
    rc = lib.jamun_tail_pack(fake, None, fake, 9, 9, 1.0, 1.0, 4, 128, fake, fake, 0, None, None, None, 0.0, 1, None)
    assert rc != 0 and b"unsupported input irreps" in lib.jamun_last_error()
    # empty problems are accepted without a launch
    assert lib.jamun_conv_p2(fake, fake, fake, fake, fake, fake, 0, fake, fake, 96, 1.0, None, None) == 0
    assert lib.jamun_tail_mix(fake, None, None, None, 0, fake, None, None, 0, None) == 0


def test_walk_params_struct_matches_header():
    from jamun_b200 import _lib

    header = open(os.path.join(ROOT, "include", "jamun_b200.h")).read()
    end = header.index("} jamun_walk_params;")
    body = header[header.rindex("typedef struct {", 0, end):end]
    names = re.findall(r"\b([a-z_0-9]+)\s*[,;]", body)
    assert names == [f[0] for f in _lib.WalkParams._fields_]
    # the fused-epilogue descriptor of jamun_gemm_f16x3_fused
    end = header.index("} jamun_gemm_epilogue;")
    body = re.sub(r"/\*.*?\*/", "", header[header.rindex("typedef struct {", 0, end):end], flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*[,;]", body)
    assert names == [f[0] for f in _lib.GemmEpilogue._fields_]


def test_no_cpu_fallback_and_error_conventions():
    import jamun_b200 as J
    from jamun_b200 import data, ops
    from jamun_b200.sampling.mcmc import ABOBA, BAOAB

    with pytest.raises(RuntimeError):
        ops.center_scale(torch.zeros(3, 3), torch.tensor([0, 3], dtype=torch.int32), 1.0)
    with pytest.raises(ValueError):
        J.default_denoiser(add_fixed_noise=True, add_fixed_ones=True)
    for cls in (BAOAB, ABOBA):
        with pytest.raises(RuntimeError):
            cls(v_init="bogus")
    m = J.default_denoiser()
    b = data.Batch.from_tensors(synthetic.make_tensors([5, 4]))
    with pytest.raises(RuntimeError):  # CPU tensors are rejected, never silently computed
        m.xhat(b, 0.04)
    with pytest.raises(NotImplementedError):
        J.e3tools.nn.ConvBlock("10x0e+4x1e", "16x0e+4x1e", "1x0e+1x1e", 64).pack(torch.zeros(2, 32))


def test_product_never_imports_oracle():
    import subprocess
    import sys

    code = "import sys, jamun_b200, jamun_b200.sampling, jamun_b200.model; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jamun_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_state_dict_matches_reference_layout(models):
    import jamun_b200 as J

    o32, _, prod = models
    a = {k: tuple(v.shape) for k, v in o32.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in prod.state_dict().items()}
    assert a == b
    compiled = J.default_denoiser(use_torch_compile=True)
    assert all(k.startswith("g._orig_mod.") for k in compiled.state_dict())
    compiled.load_state_dict(o32.state_dict())  # g. -> g._orig_mod.
    sd = dict(compiled.state_dict())
    sd["g._orig_mod.layers.0.gated_conv.f.f.tp.weight"] = torch.zeros(0)  # e3nn constant buffers are tolerated
    sd["g._orig_mod.layers.0.gated_conv.f.f.tp.output_mask"] = torch.ones(248)
    sd["g._orig_mod.layers.0.gated_conv.f.f.tp._compiled_main_left_right._w3j_1_1_1"] = torch.zeros(3, 3, 3)
    plain = J.default_denoiser(use_torch_compile=False)
    plain.load_state_dict(sd)
    assert torch.equal(plain.state_dict()["g.output_head.1.weight"], o32.state_dict()["g.output_head.1.weight"])


def test_checkpoint_roundtrip(tmp_path, models):
    import jamun_b200 as J

    o32 = models[0]
    m = J.default_denoiser(use_torch_compile=True)
    m.load_state_dict(o32.state_dict())
    path = tmp_path / "last.ckpt"
    torch.save({"state_dict": m.state_dict(), "hyper_parameters": m.hparams}, path)
    m2 = J.model.Denoiser.load_from_checkpoint(str(path))
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_packed_operands_reproduce_oracle_fp64(models):
    """Aggregate-then-transform algebra + Conv.pack / Linear.packed re-layout == the reference formulation (fp64)."""
    import jamun_b200 as J

    o32, o64, _ = models
    prod = J.default_denoiser()
    prod.load_state_dict(o32.state_dict())
    g = prod.double().arch_module
    t = synthetic.make_tensors([22, 15, 9, 30])
    b = make_oracle_batch(t, torch.float64)
    N = b.pos.shape[0]
    sig = torch.tensor(0.04, dtype=torch.float64)
    gen = torch.Generator().manual_seed(0)
    ybar = O.mean_center_pos(b.pos + 0.04 * torch.randn(N, 3, generator=gen, dtype=torch.float64), b.batch, 4)
    c_in, _, _, c_noise = o64.normalization_factors(sig, 0.332)
    r_cut = o64.effective_radial_cutoff(sig) / c_in
    yg = o64.add_edges(b.with_pos(ybar), r_cut)
    ys = yg.with_pos(ybar * c_in)
    with torch.no_grad():
        gref, hidden = o64.g(ys, c_noise.unsqueeze(0), r_cut, return_hidden=True)
        src, dst = yg.edge_index
        order = torch.sort(dst, stable=True).indices
        src, dst, eb = src[order], dst[order], yg.bond_mask[order]
        rhat, rb = KM.edge_geom(ys.pos, src, dst, r_cut)
        emb = g.embed_bondedness.weight
        blocks = [g.initial_projector.pack(emb)] + [l.pack(emb) for l in g.layers]
        cn = float(c_noise)
        s_init = KM.noise_mlp(*g.initial_noise_scaling.mlp_operands(), cn, False)
        scales = [KM.noise_mlp(*mm.mlp_operands(), cn, False) for mm in g.noise_scalings]
        skips = [KM.noise_mlp(*mm.weights.mlp_operands(), cn, True) for mm in g.skip_connections]
        x_in, x_res = o64.g.atom_embedder(b) * s_init, None
        for l, bl in enumerate(blocks):
            h = KM.radial_hidden(rb, eb, bl["w0r"], bl["b0eff"])
            co = KM.conv(x_in, bl["s_in"], bl["v_in"], src, dst, h, rhat, bl["m0"], bl["m1"], bl["alpha0"], bl["alpha1"], N)
            x_new, x_sc = KM.block_tail(co, x_in, bl["s_in"], bl["v_in"], x_res, bl, skips[l - 1] if l > 0 else None,
                                        scales[l] if l < 5 else None)
            assert torch.allclose(KM.from_soa(x_new, 120, 32), hidden[l], atol=1e-12), l
            x_in, x_res = x_sc, x_new
        hb, lin2 = g.output_head[0], g.output_head[1]
        gg = KM.head(x_res, hb.lin.packed(0), hb.lin.packed(1), lin2.packed(1).reshape(-1) * g.output_gain, hb.gate.c_gate)
    assert torch.allclose(gg, gref, atol=1e-13)
    assert blocks[1]["m0"].shape == (65, 152, 152) and blocks[1]["m1"].shape == (65, 184, 32)
    assert blocks[0]["m0"].shape == (65, 56, 152) and blocks[0]["m1"].shape == (65, 56, 32)


def test_gate_constants_and_irreps():
    from jamun_b200.e3tools.nn import _gate
    from jamun_b200.irreps import Irreps

    assert _gate.normalize2mom_const(lambda z: torch.nn.functional.leaky_relu(z, 0.01)) == pytest.approx(_gate.C_LEAKY_RELU, rel=1e-12)
    assert _gate.normalize2mom_const(torch.sigmoid) == pytest.approx(_gate.C_SIGMOID, rel=1e-12)
    ir = Irreps("120x0e + 32x1e")
    assert ir.dim == 216 and ir.num_irreps == 152 and ir.scalars_vectors() == (120, 32)
    assert Irreps("8x0e+8x0e+32x0e+8x0e").scalars_vectors() == (56, 0)
    assert repr(_gate.Gate(ir).irreps_in) == "152x0e+32x1e"
    with pytest.raises(NotImplementedError):
        Irreps("4x2e").scalars_vectors()


def test_batch_container_and_unbatch():
    from jamun_b200 import data, utils

    chains = [synthetic.make_chain(n, i) for i, n in enumerate([5, 3, 4])]
    ds = [data.DataWithResidueInformation(**{k: torch.as_tensor(v) for k, v in c.items()}) for c in chains]
    b = data.Batch.from_data_list(ds)
    assert b.num_graphs == 3 and b.num_nodes == 12 and b.batch.tolist() == [0] * 5 + [1] * 3 + [2] * 4
    assert b.edge_index.shape[1] == 4 + 2 + 3 and int(b.edge_index[:, 4:6].min()) >= 5
    c = b.clone("pos")
    c.pos += 1
    assert not torch.equal(c.pos, b.pos) and c.atom_type_index is b.atom_type_index

    class Dummy:
        device = torch.device("cpu")

    w = utils.ModelSamplingWrapper(Dummy(), b, 0.04)
    out = w.unbatch_samples({"xhat": torch.arange(36.0).reshape(12, 3), "xhat_traj": torch.zeros(7, 12, 3), "t": torch.ones(7)})
    assert [o["xhat"].shape[0] for o in out] == [5, 3, 4] and out[1]["xhat_traj"].shape == (3, 7, 3)
    with pytest.raises(AssertionError):
        w.positions_to_graph(torch.zeros(11, 3))


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jamun_b200.sampling import Sampler
    from jamun_b200.sharding import shard_chains

    sizes = synthetic.workload_sizes("2AA", 10)
    mine = shard_chains(len(sizes), rank, world)
    sample = torch.full((sum(sizes[i] for i in mine), 3), float(rank))
    allx = Sampler.gather_samples_ragged(sample)
    q.put((rank, list(mine), allx.shape[0], float(allx.sum())))
    dist.destroy_process_group()


def test_chain_sharding_and_gather_world2():
    """N>1 path on CPU: contiguous chain shards + the single final gather, over gloo with world_size 2."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    sizes = synthetic.workload_sizes("2AA", 10)
    assert res[0][1] == list(range(0, 5)) and res[1][1] == list(range(5, 10))
    total = sum(sizes)
    assert res[0][2] == res[1][2] == total
    assert res[0][3] == pytest.approx(3.0 * sum(sizes[5:]))


def test_jamun_alias_resolves_reference_targets():
    """Hydra `_target_` strings / pickled hyper-parameters of the reference resolve to this implementation."""
    import importlib
    import pickle

    import jamun  # noqa: F401
    import jamun_b200

    targets = {"jamun.model.Denoiser": jamun_b200.model.Denoiser, "jamun.model.arch.E3Conv": jamun_b200.model.arch.E3Conv,
               "jamun.e3tools.nn.ConvBlock": jamun_b200.e3tools.nn.ConvBlock, "jamun.e3tools.nn.Conv": jamun_b200.e3tools.nn.Conv,
               "jamun.e3tools.nn.EquivariantMLP": jamun_b200.e3tools.nn.EquivariantMLP,
               "jamun.sampling.Sampler": jamun_b200.sampling.Sampler,
               "jamun.sampling.walkjump.SingleMeasurementSampler": jamun_b200.sampling.walkjump.SingleMeasurementSampler,
               "jamun.sampling.mcmc.BAOAB": jamun_b200.sampling.mcmc.BAOAB, "jamun.sampling.mcmc.ABOBA": jamun_b200.sampling.mcmc.ABOBA,
               "jamun.utils.ModelSamplingWrapper": jamun_b200.utils.ModelSamplingWrapper,
               "jamun.distributions.ConstantSigma": jamun_b200.distributions.ConstantSigma}
    for path, obj in targets.items():
        mod, name = path.rsplit(".", 1)
        assert getattr(importlib.import_module(mod), name) is obj, path
    assert pickle.loads(pickle.dumps(jamun_b200.default_arch())).func is jamun_b200.model.arch.E3Conv


def test_training_conv_formulation_matches_kernel_model():
    """tests/torch_reference.py (the autograd reference the GPU tests hold the backward kernels to) evaluates the conv as a
    degree-padded batched GEMM; its algebra is checked here against the per-edge reference emulation."""
    import kernel_model as KM
    import torch_reference as train

    gen = torch.Generator().manual_seed(11)
    N, E = 13, 70
    dst = torch.sort(torch.randint(0, N - 1, (E,), generator=gen)).values  # node N-1 stays isolated
    src = torch.randint(0, N, (E,), generator=gen)
    rowptr = torch.zeros(N + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=N), 0)
    rhat = torch.nn.functional.normalize(torch.randn(E, 3, generator=gen, dtype=torch.float64), dim=1)
    h = torch.randn(E, 64, generator=gen, dtype=torch.float64)
    eid_pad, deg = train._padded_edge_index(rowptr, E)
    assert int(deg[-1]) == 0 and eid_pad.shape[0] == N
    for s_in, v_in in ((56, 0), (120, 32)):
        x = torch.randn(N, s_in + 3 * v_in, generator=gen, dtype=torch.float64)
        pk = dict(m0=torch.randn(65, s_in + v_in, 152, generator=gen, dtype=torch.float64),
                  m1=torch.randn(65, s_in + 2 * v_in, 32, generator=gen, dtype=torch.float64), alpha0=0.7, alpha1=1.3)
        got = train._conv(x, s_in, v_in, src, eid_pad, deg, h, rhat, pk)
        want = KM.conv(x, s_in, v_in, src, dst, h, rhat, pk["m0"], pk["m1"], 0.7, 1.3, N)
        assert torch.allclose(got, want, rtol=1e-10, atol=1e-10)


import warnings


def test_load_state_dict_warns_on_unknown_tensor_product_entries(models):
    import jamun_b200 as J

    o32, _, _ = models
    sd = dict(o32.state_dict())
    key = next(k for k in sd if "radial_nn.3.weight" in k).replace("radial_nn.3.weight", "tp.mystery")
    sd[key] = torch.ones(5)
    sd[key.replace("mystery", "weight")] = torch.zeros(0)  # e3nn's empty external-weight buffer: silently ignored
    m = J.default_denoiser()
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        m.load_state_dict(sd)
    assert any("tp.mystery" in str(w.message) for w in rec) and not any("tp.weight" in str(w.message) for w in rec)


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist

    from jamun_b200.ddp import GradientReducer, broadcast_parameters

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)  # replicas start different on purpose: broadcast_parameters must fix that
    model = torch.nn.Sequential(torch.nn.Linear(6, 300), torch.nn.Tanh(), torch.nn.Linear(300, 5), torch.nn.Linear(5, 1))
    broadcast_parameters(model)
    red = GradientReducer(model, bucket_bytes=1024)  # several buckets; the 300x6 weight is a bucket of its own
    data = [torch.randn(8, 6, generator=torch.Generator().manual_seed(7 + r)) for r in range(world)]
    # reference: the mean over ranks of the local gradients, computed redundantly without any collective
    ref = None
    for r in range(world):
        gs = torch.autograd.grad(model(data[r]).pow(2).mean(), list(model.parameters()))
        ref = [g / world for g in gs] if ref is None else [a + g / world for a, g in zip(ref, gs)]
    out = []
    for step in range(2):  # second step: buckets re-zeroed, views still attached
        red.reset()
        model(data[rank]).pow(2).mean().backward()
        red.finish()
        out.append([p.grad.clone() for p in model.parameters()])
    ok = all(torch.allclose(g, r_, rtol=1e-5, atol=1e-7) for g, r_ in zip(out[0], ref)) and \
        all(torch.equal(a, b) for a, b in zip(out[0], out[1]))
    flat = torch.cat([g.reshape(-1) for g in out[1]])
    q.put((rank, ok, len(red.buckets), flat.tolist()))
    dist.destroy_process_group()


def test_ddp_gradient_allreduce_world2():
    """DDP row (e2): after a step every rank holds the same gradients, equal to the mean of the per-rank gradients."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] and res[1][1]
    assert res[0][2] >= 3
    assert res[0][3] == res[1][3]  # bit-identical across ranks
