"""Device-side execution plan of the denoiser: topology constants, packed weights, workspaces and the
kernel sequence of one E3Conv evaluation (arch/e3conv.py:110-138 of the reference).

Nothing here computes on the host: tensors are device buffers handed to the C ABI (jamun_b200.ops).
torch is used for allocation, one-off weight re-layout (indexing/concatenation when a plan is built)
and stream plumbing only.
"""
from __future__ import annotations

import itertools
import math
import os
from typing import Dict, List, Optional

import torch

from . import ops, packing
from .irreps import Irreps


# "tc": aggregate builder + tcgen05 3xTF32 GEMM (product path); "simt": the exact-fp32 CUDA-core kernel (kept for
# A/B validation of the tensor-core path, selectable with JAMUN_B200_CONV=simt)
CONV_IMPL = os.environ.get("JAMUN_B200_CONV", "tc")
# aggregate builder of the tensor-core path: "tc" = tcgen05 per-node products (jamun_conv_build_tc) with the 0e(x)1e->1e
# gather (jamun_conv_p2) on a side stream; "ffma" = the FP32-pipe builder jamun_conv_build_a (exact fp32 aggregate)
BUILD_IMPL = os.environ.get("JAMUN_B200_BUILD", "tc")
# radial MLP hidden layer: "ffma" = packed-FP32 CUDA-core kernel (0.186 ms at the 2AA bench size); "mma" = warp-level tensor cores
# (tf32 three-product split, 34 % fewer instructions but 0.207 ms: latency-bound on its weight-fragment loads -- DESIGN.md 5)
RADIAL_IMPL = os.environ.get("JAMUN_B200_RADIAL_HIDDEN", "ffma")
# tensor-core split of the sampling path's GEMMs: "f16" = fp16 hi/lo split, tcgen05 kind::f16 (jamun_gemm_f16x3: half the MMA
# instructions of the tf32 form at the same 11-bit significands; weights pre-scaled into the fp16 range when the plan is built,
# operand overflow reported through Topology.gemm_status); "tf32" = jamun_gemm_tf32x3.  Read when a plan is built.
GEMM_KIND = os.environ.get("JAMUN_B200_GEMM", "f16")


def fall_back_to_tf32(what: str) -> None:
    """An activation left the fp16 range in the fp16-split GEMMs: switch this process to the tf32-split kernels (fp32 range,
    same accuracy, ~1.4x slower contraction) so the caller can redo the call.  Still this library's tensor-core kernels --
    not a CPU or library fallback."""
    import warnings

    warnings.warn(f"jamun_b200: an activation exceeded the fp16 range (65504) during {what}; switching to the tf32-split "
                  "tensor-core GEMMs (JAMUN_B200_GEMM=tf32) and repeating the call", RuntimeWarning, stacklevel=3)
    os.environ["JAMUN_B200_GEMM"] = "tf32"
# layout of the conv operand between the tensor-core builder and the fp16-split contraction: "tile" = [row/128][stage][128][32]
# (a tile's stages contiguous), "stage" = [stage][rows_pad][32] (the layout of every other user of the GEMM).  Read at call time.
A_LAYOUT = os.environ.get("JAMUN_B200_A_LAYOUT", "tile")
CELL_LIST_MIN_CHAIN = 2560  # longer chains use the cell-list search.  Measured on B200 (tools/time_radius.py, 512 k atoms, CSR build incl. the out-edge index): n=1000 brute 2.37 ms / cells 2.99 ms; n=3000 3.89 / 3.75; n=6000 6.21 / 5.59 -- the ascending scan stops after 33 hits, so the O(n^2) bound only bites beyond ~2.5 k atoms
Y_LD = 17 * 128  # row stride of the per-node transform Y (65*32 = 2080 columns padded to 17 column blocks of 128)


class Topology:
    WORKSPACE_BYTES = 24 << 30

    """Per-template constants resident in HBM: chain layout, bonded CSR by receiver, atom indices, and the
    edge workspaces sized for the worst case so that no kernel ever needs a host-side edge count."""

    def __init__(self, batch, device, max_num_neighbors: Optional[int] = 32):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("jamun_b200 runs on CUDA devices only (no CPU fallback)")
        bvec = batch["batch"] if "batch" in batch else torch.zeros(batch.num_nodes, dtype=torch.long)
        bvec = bvec.detach().to("cpu", torch.long)
        N = int(bvec.numel())
        G = int(bvec.max().item()) + 1 if N else 0
        if N and not bool((bvec[1:] >= bvec[:-1]).all()):
            raise ValueError("batch vector must be sorted (PyG Batch layout)")
        counts = torch.bincount(bvec, minlength=G)
        ptr = torch.zeros(G + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(counts, 0)
        self.N, self.G = N, G
        self.max_chain = int(counts.max().item()) if N else 0
        self.device = dev
        self.max_num_neighbors = max_num_neighbors
        self.chain_ptr = ptr.to(dev, torch.int32)
        self.chain_ptr_long = ptr.to(dev)
        self.chain_of = bvec.to(dev, torch.int32)
        self.batch_long = bvec.to(dev)
        ei = batch["edge_index"].detach().to("cpu", torch.long) if "edge_index" in batch else torch.zeros(2, 0, dtype=torch.long)
        order = torch.sort(ei[1], stable=True).indices  # receiver-major, original order within a receiver
        self.bond_src = ei[0][order].to(dev, torch.int32).contiguous()
        brow = torch.zeros(N + 1, dtype=torch.long)
        brow[1:] = torch.cumsum(torch.bincount(ei[1], minlength=N), 0)
        self.bond_rowptr = brow.to(dev, torch.int32)
        self.num_bonded = int(ei.shape[1])
        if self.bond_src.numel() == 0:
            self.bond_src = torch.zeros(1, dtype=torch.int32, device=dev)
        if max_num_neighbors is None:
            cap = int((counts * (counts - 1)).sum().item()) + self.num_bonded
        else:
            per = torch.minimum(counts - 1, torch.full_like(counts, max_num_neighbors + 1)).clamp_min(0)
            cap = int((counts * per).sum().item()) + self.num_bonded
        self.cap = max(cap, 1)
        bond_in = int(torch.bincount(ei[1], minlength=max(N, 1)).max().item()) if ei.shape[1] else 0
        nbr = int(counts.max().item()) - 1 if N else 0
        self.max_degree = (nbr if max_num_neighbors is None else min(nbr, max_num_neighbors + 1)) + bond_in  # in-degree bound
        i32 = dict(dtype=torch.int32, device=dev)
        self.idx = []
        for key in ("atom_type_index", "atom_code_index", "residue_code_index", "residue_sequence_index"):
            self.idx.append(batch[key].detach().to(dev, torch.int32).contiguous() if key in batch else None)
        # CSR + edge workspaces
        self.rowptr = torch.zeros(N + 1, **i32)
        self.scratch = torch.zeros(N + 1, **i32)
        self.src_rowptr = torch.zeros(N + 1, **i32)  # out-edge lists (CSR by source) for the source-major path-2 pass
        self.src_eid = torch.zeros(self.cap, **i32)
        self.col = torch.zeros(self.cap, **i32)
        self.edst = torch.zeros(self.cap, **i32)
        self.ebond = torch.zeros(self.cap, dtype=torch.uint8, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.rhat = torch.zeros(self.cap, 4, **f32)
        self.rb = torch.zeros(self.cap, ops.NBASIS, **f32)
        self.h = torch.zeros(self.cap, ops.EDGE_HID, **f32)
        # node workspaces
        self.x0 = torch.zeros(N, ops.S0, **f32)
        self.xa = [torch.zeros(N, ops.HID, **f32) for _ in range(2)]  # node_attr ping-pong
        self.xs = [torch.zeros(N, ops.HID, **f32) for _ in range(2)]  # scaled node_attr ping-pong
        self.conv = torch.zeros(N, ops.GATE_IN, **f32)
        self.inv_deg = torch.zeros(max(N, 1), **f32)
        self.x0_key = None
        # fp32 A operand of the tensor-core conv (stage-major), sized for the hidden blocks; large batches are
        # processed in row chunks so the workspace stays below WORKSPACE_BYTES
        per_row = 65 * (5 + 3 * 2) * 32 * 4
        rows_pad = (N + 127) // 128 * 128
        max_rows = max(128, (self.WORKSPACE_BYTES // per_row) // 128 * 128)
        self.chunk_rows = min(rows_pad, max_rows)
        self.a_ws = None  # allocated on first use
        self.side_stream = torch.cuda.Stream(device=dev)
        self.ev_fork, self.ev_join = torch.cuda.Event(), torch.cuda.Event()
        self.p2_pending = False
        self.xs_op_of = None  # the tensor whose scalars currently sit packed in xs_op (written by tail_mix)
        self.y0 = None
        self.y0_key = None
        self.csr_generation = 0  # bumped by every writer of the CSR buffers (Denoiser.add_edges tags batches with it)
        self.gemm_status = torch.zeros(1, dtype=torch.int32, device=dev)  # bit 0: an fp16-split GEMM operand left the fp16 range

    def check_status(self) -> None:
        """Host sync: raise if a kernel of the fp16-split path reported an operand outside the fp16 range since the last check."""
        if int(self.gemm_status.item()) != 0:
            self.gemm_status.zero_()
            raise FloatingPointError("jamun_b200: an activation exceeded the fp16 range (65504) in the fp16-split tensor-core GEMM; "
                                     "results of this call are invalid -- set JAMUN_B200_GEMM=tf32 and rebuild the plan")

    def overflowed(self) -> bool:
        """Host sync: True (and the flag is cleared) if an fp16-split GEMM reported an operand outside the fp16 range."""
        if int(self.gemm_status.item()) == 0:
            return False
        self.gemm_status.zero_()
        return True

    def build_csr(self, pos: torch.Tensor, r_cut: float):
        """K1 on (mean-centred, unscaled) positions; r2 = float(double(r)*double(r)) as torch_cluster does."""
        r2 = float(torch.tensor(float(r_cut) * float(r_cut), dtype=torch.float64).to(torch.float32))
        mnn = -1 if self.max_num_neighbors is None else int(self.max_num_neighbors)
        impl = os.environ.get("JAMUN_B200_RADIUS", "auto")  # "brute" | "cells" | "auto" (cell list for chains of > 64 atoms)
        cells_ok = 0 <= mnn < 64 and self.max_chain <= 8192
        if cells_ok and (impl == "cells" or (impl == "auto" and self.max_chain > CELL_LIST_MIN_CHAIN)):
            if getattr(self, "nbr_tmp", None) is None:
                self.nbr_tmp = torch.zeros(max(self.N, 1) * (mnn + 1), dtype=torch.int32, device=self.device)
            ops.radius_csr_cells(pos, self.chain_ptr, self.max_chain, r2, float(r_cut), mnn, self.bond_rowptr, self.bond_src,
                                 self.scratch, self.nbr_tmp, self.rowptr, self.col, self.edst, self.ebond)
        else:
            ops.radius_csr(pos, self.chain_of, self.chain_ptr, r2, mnn, self.bond_rowptr, self.bond_src, self.scratch,
                           self.rowptr, self.col, self.edst, self.ebond)
        ops.csr_by_source(self.rowptr, self.col, self.scratch, self.src_rowptr, self.src_eid)
        self.csr_generation += 1

    def set_csr_from_edge_index(self, edge_index: torch.Tensor, bond_mask: torch.Tensor):
        """Compatibility path for callers that hand E3Conv.forward an explicit edge list."""
        ei = edge_index.to(self.device, torch.long)
        E = ei.shape[1]
        if E > self.cap:
            raise ValueError(f"{E} edges exceed the workspace capacity {self.cap}")
        order = torch.sort(ei[1], stable=True).indices
        self.col[:E] = ei[0][order].to(torch.int32)
        self.edst[:E] = ei[1][order].to(torch.int32)
        self.ebond[:E] = bond_mask.to(self.device)[order].to(torch.uint8)
        rp = torch.zeros(self.N + 1, dtype=torch.long, device=self.device)
        rp[1:] = torch.cumsum(torch.bincount(ei[1], minlength=self.N), 0)
        self.rowptr.copy_(rp.to(torch.int32))
        ops.csr_by_source(self.rowptr, self.col, self.scratch, self.src_rowptr, self.src_eid)
        self.csr_generation += 1

    def edge_index(self):
        """Materialise (edge_index [2,E] int64, bond_mask [E] int64) from the CSR -- host sync; tests/API only."""
        E = int(self.rowptr[-1].item())
        ei = torch.stack([self.col[:E].long(), self.edst[:E].long()])
        return ei, self.ebond[:E].long()


class E3ConvPlan:
    """Kernel operands of one E3Conv module at one noise level (c_noise)."""

    _serials = itertools.count(1)

    def __init__(self, g, c_noise: float, device):
        from .e3tools.nn import ConvBlock, EquivariantMLP
        from .model.atom_embedding import AtomEmbeddingWithResidueInformation

        dev = torch.device(device)
        if Irreps(g.irreps_sh) != Irreps("1x0e+1x1e") or Irreps(g.irreps_out) != Irreps("1x1e") \
                or g.edge_attr_dim != 2 * ops.NBASIS:
            raise NotImplementedError("B200 kernels are built for irreps_sh=1x0e+1x1e, irreps_out=1x1e, edge_attr_dim=64")
        if not isinstance(g.atom_embedder, AtomEmbeddingWithResidueInformation):
            raise NotImplementedError("use_residue_information=False is outside the kernels' scope")
        blocks = [g.initial_projector, *g.layers]
        if not all(isinstance(b, ConvBlock) for b in blocks) or not isinstance(g.output_head, EquivariantMLP) \
                or len(g.output_head) != 2:
            raise NotImplementedError("hidden_layer_factory must be ConvBlock and output_head_factory EquivariantMLP([hidden])")
        self.c_noise = float(c_noise)
        self.device = dev
        self.serial = next(E3ConvPlan._serials)  # cache key for plan-dependent constants held by topologies
        self.gemm_kind = os.environ.get("JAMUN_B200_GEMM", GEMM_KIND)
        if self.gemm_kind not in ("f16", "tf32"):
            raise ValueError(f"JAMUN_B200_GEMM={self.gemm_kind!r}: expected f16 or tf32")
        # Parameter re-layout happens on a host copy of the module (pure indexing, once per plan) and the results are uploaded;
        # the (hi | lo) operand images are produced on the device by jamun_pack_b -- one launch per operand, so building a
        # plan issues a few dozen launches of this library's kernels and no framework indexing kernels.
        import copy

        saved, g._plan = g._plan, None
        try:
            with torch.no_grad():  # parameters / buffers are copied device -> host directly (memo), the module tree is cloned
                memo = {id(t): torch.nn.Parameter(t.detach().cpu(), requires_grad=False) for t in g.parameters()}
                memo.update({id(t): t.detach().cpu() for t in g.buffers()})
                gc = copy.deepcopy(g, memo)
        finally:
            g._plan = saved
        up = lambda t: t.detach().to(torch.float32).contiguous().to(dev)  # noqa: E731  (host -> device copy, no kernel)
        with torch.no_grad():
            emb = gc.embed_bondedness.weight.detach().to(torch.float32)
            self.tables = [up(t) for t in gc.atom_embedder.tables()]
            if sum(t.shape[1] for t in self.tables) != ops.S0:
                raise NotImplementedError("atom embedding width must be 56")
            self.use_residue_sequence_index = g.atom_embedder.use_residue_sequence_index
            self.blocks: List[Dict] = []
            for b in [gc.initial_projector, *gc.layers]:
                pk = b.pack(emb)
                host = {k: (v.detach().to(torch.float32).contiguous() if isinstance(v, torch.Tensor) else v) for k, v in pk.items()}
                ns_in = (host["s_in"] + 31) // 32
                # block tail as one GEMM: [activated scalars (128) | input scalars (32 NS)] . [W_self ; W_skip], same for vectors
                ws = torch.zeros(128 + 32 * ns_in, 120, dtype=torch.float32)
                ws[:120] = host["wself_s"]
                ws[128:128 + host["s_in"]] = host["wskip_s"]
                wv = host["wself_v"] if host["wskip_v"] is None else torch.cat([host["wself_v"], host["wskip_v"]], dim=0)
                blk = {k: (up(v) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
                blk["gemm_kind"] = self.gemm_kind
                if self.gemm_kind == "f16":
                    # power-of-two pre-scales from the host copies (no device sync); undone through the GEMMs' alpha
                    sc = blk["f16_scales"] = (ops.f16_scale(host["m0"]), ops.f16_scale(host["m1"]), ops.f16_scale(ws), ops.f16_scale(wv))
                    blk["b0_img"], blk["b1_img"], blk["wy_img"] = packing.pack_conv_operands_device(
                        blk["m0"], blk["m1"], blk["s_in"], blk["v_in"], f16_scales=sc[:2])
                    blk["tail_bs_img"] = ops.pack_b_f16(up(ws), ws.shape[0] // 32, 128, sc[2])
                    blk["tail_bv_img"] = ops.pack_b_f16(up(wv), wv.shape[0] // 32, 32, sc[3])
                else:
                    blk["f16_scales"] = (1.0, 1.0, 1.0, 1.0)
                    blk["b0_img"], blk["b1_img"], blk["wy_img"] = packing.pack_conv_operands_device(blk["m0"], blk["m1"], blk["s_in"],
                                                                                                    blk["v_in"])
                    blk["tail_bs_img"] = ops.pack_b(up(ws), n_stages=ws.shape[0] // 32, n_pad=128)
                    blk["tail_bv_img"] = ops.pack_b(up(wv), n_stages=wv.shape[0] // 32, n_pad=32)
                self.blocks.append(blk)
            self.w0r_all = up(torch.stack([b.pack(emb)["w0r"] for b in [gc.initial_projector, *gc.layers]]))      # [L, 32, 64]
            self.b0eff_all = up(torch.stack([b.pack(emb)["b0eff"] for b in [gc.initial_projector, *gc.layers]]))  # [L, 2, 64]
            self.w0r_frag = ops.radial_pack_frag(self.w0r_all) if RADIAL_IMPL == "mma" else None  # tensor-core weight images
            self.set_noise(g, self.c_noise, bump=False)
            blk, lin2 = gc.output_head[0], gc.output_head[1]
            self.head_w1s = up(blk.lin.packed(0))
            self.head_w1v = up(blk.lin.packed(1))
            self.head_cgate = blk.gate.c_gate
            self.head_w2 = up(lin2.packed(1).reshape(-1) * gc.output_gain.detach())
        self.n_basis = ops.NBASIS
        self._grids: Dict[float, tuple] = {}  # r_cut -> (centres on the device, spacing); never evicted (see radial_grid)

    def set_noise(self, g, c_noise: float, bump: bool = True) -> None:
        """The only noise-level-dependent operands: the outputs of the 11 noise-conditioning MLPs (jamun_noise_mlp).  A new noise
        level on an existing plan recomputes just these (the packed weight images are sigma-independent) into *fresh* tensors and
        takes a new serial, so CUDA graphs and per-topology constants captured for the previous level are never reused."""
        dev = self.device
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()  # noqa: E731
        self.c_noise = float(c_noise)
        with torch.no_grad():
            self.s_init = ops.noise_mlp(*map(f32, g.initial_noise_scaling.mlp_operands()), self.c_noise, False)
            self.scales = [ops.noise_mlp(*map(f32, m.mlp_operands()), self.c_noise, False) for m in g.noise_scalings]
            self.skips = [ops.noise_mlp(*map(f32, m.weights.mlp_operands()), self.c_noise, True) for m in g.skip_connections]
        if bump:
            self.serial = next(E3ConvPlan._serials)

    def radial_grid(self, r_cut: float):
        """soft_one_hot_linspace(..., cutoff=True) grid: centres linspace(0, r, n+2)[1:-1] and their spacing.

        The device tensor is cached on the plan (keyed by r_cut): CUDA graphs captured by fused_baoab bake its raw pointer
        into edge_geom through the C ABI, so it must outlive every graph that replays it."""
        key = float(r_cut)
        hit = self._grids.get(key)
        if hit is None:
            values = torch.linspace(0.0, key, self.n_basis + 2, dtype=torch.float32)
            step = float(values[1] - values[0])
            hit = self._grids[key] = (values[1:-1].to(self.device).contiguous(), step)
        return hit


def _gemm(topo: Topology, kind: str, a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, rows: int, rows_pad: int, rs_ptr, out_ptr,
          out_ld: int, **kw) -> None:
    """One launch of the node-tile GEMM in the plan's operand format (alpha already divided by the fp16 weight pre-scale)."""
    if kind == "f16":
        ops.gemm_f16x3(a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, rows, rows_pad, rs_ptr, out_ptr, out_ld,
                       status=topo.gemm_status, **kw)
    else:
        kw.pop("addend_scale", None)
        assert not kw.pop("a_tile_major", False), "the tile-major operand layout needs the fp16-split GEMM"
        ops.gemm_tf32x3(a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, rows, rows_pad, rs_ptr, out_ptr, out_ld, **kw)


def _contract(topo: Topology, a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, nrows: int, rp: int, rs_ptr, out_ptr,
              addend=None, kind: str = "tf32", addend_scale=None, tiled: bool = False) -> None:
    """The contraction GEMM; with few row tiles (small batches) the K stages are split over several CTAs per tile."""
    tiles = (nrows + 127) // 128
    if addend is not None:
        ks = 1
    elif tiles <= 74:  # few tiles: fill the SMs
        ks = min(16, 148 // tiles, min(n_stages))
    else:              # (balancing the last wave of multi-wave launches with a 2-4x split was measured: no gain)
        ks = 1
    if os.environ.get("JAMUN_B200_SPLITK", "1") != "1":
        ks = 1
    if ks > 1:
        need = ks * nrows * ops.GATE_IN
        if getattr(topo, "gemm_partial", None) is None or topo.gemm_partial.numel() < need:
            topo.gemm_partial = torch.empty(need, dtype=torch.float32, device=topo.device)
        if kind == "f16":
            ops.gemm_f16x3(a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, nrows, rp, rs_ptr, out_ptr, ops.GATE_IN, k_splits=ks,
                           partial=topo.gemm_partial, status=topo.gemm_status, a_tile_major=tiled)
        else:
            ops.gemm_tf32x3_splitk(a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, nrows, rp, rs_ptr, out_ptr, ops.GATE_IN, ks,
                                   topo.gemm_partial)
    else:
        _gemm(topo, kind, a_ptrs, b_ptrs, n_stages, n_pad, n_valid, out_col, alpha, nrows, rp, rs_ptr, out_ptr, ops.GATE_IN,
              addend_ptrs=addend, addend_ld=None if addend is None else [0, 96, 96, 96],
              addend_scale=None if addend is None else addend_scale, a_tile_major=tiled)


def _ensure_tail_operands(topo: Topology) -> None:
    """Block-tail operands (stage-major, rows_all rows): tail_s = [activated scalars (4 stages) | input scalars (4 stages)],
    tail_v = 3 components x [gated vectors | input vectors].  The input-scalar half doubles as the per-node transform's operand
    xs_op.  Written by the fused GEMM epilogues (jamun_gemm_f16x3_fused), zero where never written."""
    if getattr(topo, "tail_s", None) is None:
        rows_all = (topo.N + 127) // 128 * 128
        topo.tail_s = torch.zeros(8 * rows_all * 32, dtype=torch.float32, device=topo.device)
        topo.tail_v = torch.zeros(3 * 2 * rows_all * 32, dtype=torch.float32, device=topo.device)
        topo.xs_op = topo.tail_s[4 * rows_all * 32:]


def _fuse_gate_ok(topo: Topology, b: Dict, build_impl: str) -> bool:
    """The contraction can apply the Gate and write the block-tail operands itself (jamun_gemm_f16x3_fused mode 1) when it
    runs as one pass per row tile: fp16-split GEMM, tensor-core builder, hidden block, enough tiles that split-K is off."""
    tiles = (min(topo.chunk_rows, topo.N) + 127) // 128
    return (b.get("gemm_kind") == "f16" and build_impl == "tc" and b["v_in"] > 0 and tiles > 74
            and os.environ.get("JAMUN_B200_TAIL", TAIL_IMPL) == "tc" and os.environ.get("JAMUN_B200_TAIL_FUSE", "1") == "1")


def conv_tc(topo: Topology, b: Dict, x: torch.Tensor, out: torch.Tensor, y_const_key=None, defer_reduce: bool = False,
            fuse_gate: bool = False) -> None:
    """Conv.forward on the tensor cores: per-node transform Y = x_s.W of the 0e(x)1e->1e path (tcgen05 GEMM,
    17 column blocks) -> aggregate A of the other paths (jamun_conv_build_tc: per-node tcgen05 products; or the FP32-pipe
    jamun_conv_build_a) -> contraction jamun_gemm_tf32x3.  With the tensor-core builder the gather of Y (jamun_conv_p2) runs
    on a side stream: call conv_tc_join() before consuming `out` and add its result to out[:, 152:] (block_tail does)."""
    s_in, v_in = b["s_in"], b["v_in"]
    ns = (s_in + 31) // 32
    nsl0, nsl1 = ns + (1 if v_in else 0), (2 if v_in else 0)
    st0, st1 = 65 * nsl0, 65 * nsl1
    rp = topo.chunk_rows
    N = topo.N
    rows_all = (N + 127) // 128 * 128
    if topo.a_ws is None:
        topo.a_ws = torch.empty(65 * (5 + 3 * 2) * rp * 32, dtype=torch.float32, device=topo.device)
        _ensure_tail_operands(topo)
        topo.y = torch.empty(N, Y_LD, dtype=torch.float32, device=topo.device)
        topo.p2 = torch.empty(N, 96, dtype=torch.float32, device=topo.device)
        topo.t_edge = torch.empty(topo.cap, 32, dtype=torch.float32, device=topo.device)
    kind, sc = b.get("gemm_kind", "tf32"), b.get("f16_scales", (1.0, 1.0, 1.0, 1.0))
    wy_block = ns * 128 * 32 * (1 if kind == "f16" else 2)  # 4-byte words per column block of the wy images
    # per-node transform of the scalar inputs.  For the initial block the input (atom embedding x noise scale) does not depend
    # on positions, so its transform is computed once per (topology, plan) and kept.
    y_buf = topo.y
    if y_const_key is not None:
        if topo.y0 is None or topo.y0_key != y_const_key:
            topo.y0 = torch.empty(N, Y_LD, dtype=torch.float32, device=topo.device)
            ops.pack_rows(x, 0, s_in, rows_all, topo.xs_op)
            _gemm(topo, kind, [topo.xs_op.data_ptr()], [b["wy_img"].data_ptr()], [ns], [128], [128], [0], [1.0 / sc[1]], N, rows_all,
                  None, topo.y0.data_ptr(), Y_LD, col_blocks=17, b_block_floats=wy_block)
            topo.y0_key = y_const_key
        y_buf = topo.y0
    else:
        if topo.xs_op_of is not x:  # else the previous block's tail_mix has already written the packed scalars of x
            ops.pack_rows(x, 0, s_in, rows_all, topo.xs_op)
        topo.xs_op_of = None
        _gemm(topo, kind, [topo.xs_op.data_ptr()], [b["wy_img"].data_ptr()], [ns], [128], [128], [0], [1.0 / sc[1]], N, rows_all, None,
              topo.y.data_ptr(), Y_LD, col_blocks=17, b_block_floats=wy_block)
    base = topo.a_ws.data_ptr()
    a1_off = st0 * rp * 32
    comp = st1 * rp * 32
    build_impl = os.environ.get("JAMUN_B200_BUILD", BUILD_IMPL)
    tiled = build_impl == "tc" and kind == "f16" and os.environ.get("JAMUN_B200_A_LAYOUT", A_LAYOUT) == "tile"
    if build_impl == "tc":
        # the 0e(x)1e->1e gather needs only Y and h: it runs on the side stream under the builder and the contraction, and
        # its result joins in block_tail (hidden blocks: topo.p2 as the `vadd` operand; initial block: written in place)
        main = torch.cuda.current_stream()
        side = main if os.environ.get("JAMUN_B200_P2_STREAM", "side") == "main" else topo.side_stream
        topo.ev_fork.record(main)
        with torch.cuda.stream(side):
            side.wait_event(topo.ev_fork)
            if fuse_gate:  # raw receiver-side sums -> the contraction's addend (its epilogue gates conv + p2 together)
                ops.conv_p2(topo.rowptr, topo.src_rowptr, topo.src_eid, topo.h, topo.rhat, y_buf, topo.t_edge, topo.p2.data_ptr(), 96, 0.0)
            elif defer_reduce:  # only T_e; the receiver-side sum is taken by jamun_tail_pack (block_tail)
                ops.conv_p2(topo.rowptr, topo.src_rowptr, topo.src_eid, topo.h, topo.rhat, y_buf, topo.t_edge, None, 0, 0.0)
            elif v_in:
                ops.conv_p2(topo.rowptr, topo.src_rowptr, topo.src_eid, topo.h, topo.rhat, y_buf, topo.t_edge, topo.p2.data_ptr(), 96,
                            b["alpha1"])
            else:
                ops.conv_p2(topo.rowptr, topo.src_rowptr, topo.src_eid, topo.h, topo.rhat, y_buf, topo.t_edge,
                            out.data_ptr() + 4 * 152, ops.GATE_IN, b["alpha1"])
            topo.ev_join.record(side)
        topo.p2_pending = True
        topo.p2_deferred = bool(defer_reduce) and not fuse_gate
        topo.gate_fused = bool(fuse_gate)
    for row0 in range(0, N, rp):
        nrows = min(rp, N - row0)
        if build_impl == "tc":
            ops.conv_build_tc(x, s_in, v_in, topo.rowptr, topo.col, topo.h, topo.rhat, row0, nrows, rp, base,
                              base + 4 * a1_off if v_in else None, comp, topo.inv_deg, tiled=tiled)
        if v_in:
            if build_impl != "tc":
                ops.conv_build_a(x, s_in, v_in, topo.rowptr, topo.col, topo.h, topo.rhat, y_buf, topo.max_degree, row0, nrows, rp, base,
                                 base + 4 * a1_off, comp, topo.p2.data_ptr(), 96, 0.0, topo.inv_deg)
            a_ptrs = [base] + [base + 4 * (a1_off + c * comp) for c in range(3)]
            b_ptrs = [b["b0_img"].data_ptr()] + [b["b1_img"].data_ptr()] * 3
            p2 = topo.p2.data_ptr() + 4 * row0 * 96
            if fuse_gate:
                if topo.p2_pending:  # the receiver-side sums must be complete before the contraction's epilogue adds them
                    torch.cuda.current_stream().wait_event(topo.ev_join)
                    topo.p2_pending = False
                epi = dict(mode=1, op_s=topo.tail_s.data_ptr() + 4 * row0 * 32, op_v=topo.tail_v.data_ptr() + 4 * row0 * 32,
                           op_v_comp_stride=2 * rows_all * 32, op_rows_pad=rows_all, c_act=b["c_act"], c_gate=b["c_gate"])
                ops.gemm_f16x3_fused(a_ptrs, b_ptrs, [st0, st1, st1, st1], [160, 32, 32, 32], [152, 32, 32, 32], [0, 152, 184, 216],
                                     [b["alpha0"] / sc[0]] + [b["alpha1"] / sc[1]] * 3, nrows, rp, topo.inv_deg.data_ptr() + 4 * row0, epi,
                                     addend_ptrs=[None, p2, p2 + 4 * 32, p2 + 4 * 64], addend_ld=[0, 96, 96, 96],
                                     addend_scale=[sc[0]] + [sc[1]] * 3, status=topo.gemm_status, a_tile_major=tiled)
                continue
            addend = None if build_impl == "tc" else [None, p2, p2 + 4 * 32, p2 + 4 * 64]
            _contract(topo, a_ptrs, b_ptrs, [st0, st1, st1, st1], [160, 32, 32, 32], [152, 32, 32, 32], [0, 152, 184, 216],
                      [b["alpha0"] / sc[0]] + [b["alpha1"] / sc[1]] * 3, nrows, rp, topo.inv_deg.data_ptr() + 4 * row0,
                      out.data_ptr() + 4 * row0 * ops.GATE_IN, addend=addend, kind=kind, addend_scale=[sc[0]] + [sc[1]] * 3, tiled=tiled)
        else:  # initial block: the 1e output is the path-2 gather alone, written in place
            if build_impl != "tc":
                ops.conv_build_a(x, s_in, v_in, topo.rowptr, topo.col, topo.h, topo.rhat, y_buf, topo.max_degree, row0, nrows, rp, base, None, 0,
                                 out.data_ptr() + 4 * 152, ops.GATE_IN, b["alpha1"], topo.inv_deg)
            _contract(topo, [base], [b["b0_img"].data_ptr()], [st0], [160], [152], [0], [b["alpha0"] / sc[0]], nrows, rp,
                      topo.inv_deg.data_ptr() + 4 * row0, out.data_ptr() + 4 * row0 * ops.GATE_IN, kind=kind, tiled=tiled)


def conv_tc_join(topo: Topology, b: Dict) -> Optional[torch.Tensor]:
    """Join the side-stream gather started by conv_tc; returns the [N, 96] addend for block_tail (hidden blocks) or None."""
    if getattr(topo, "gate_fused", False):
        topo.gate_fused = False
        return "fused"  # Gate applied and operands written by the contraction's epilogue
    if not getattr(topo, "p2_pending", False):
        return None
    torch.cuda.current_stream().wait_event(topo.ev_join)
    topo.p2_pending = False
    if getattr(topo, "p2_deferred", False):
        return "deferred"  # block_tail hands t_edge to jamun_tail_pack
    return topo.p2 if b["v_in"] else None


TAIL_IMPL = os.environ.get("JAMUN_B200_TAIL", "tc")  # "tc": pack -> tcgen05 GEMM -> mix;  "simt": one exact-fp32 kernel


def block_tail(topo: Topology, b: Dict, x_in, x_res, skip_w, s_next, x_new, x_scaled, vadd) -> None:
    """Gate + self-interaction + skip Linear + noise-conditional skip/scale (ConvBlock.forward after the conv)."""
    fused = vadd == "fused" if isinstance(vadd, str) else False
    deferred = isinstance(vadd, str) and not fused
    if isinstance(vadd, str):
        vadd = None
    if os.environ.get("JAMUN_B200_TAIL", TAIL_IMPL) != "tc" or topo.a_ws is None:
        assert not deferred, "the deferred path-2 sum needs the tensor-core block tail"
        ops.block_tail(topo.conv, x_in, b["s_in"], b["v_in"], x_res, b["wself_s"], b["wself_v"], b["wskip_s"], b["wskip_v"], skip_w,
                       s_next, b["c_act"], b["c_gate"], x_new, x_scaled, vadd=vadd)
        return
    N = topo.N
    rows_all = (N + 127) // 128 * 128
    ns_in = (b["s_in"] + 31) // 32
    st_s, st_v = 4 + ns_in, 2 if b["v_in"] else 1
    need = (st_s + 3 * st_v) * rows_all * 32
    assert need <= topo.a_ws.numel(), "operand workspace too small for the block tail"
    if getattr(topo, "ytail", None) is None:
        topo.ytail = torch.empty(N, ops.HID, dtype=torch.float32, device=topo.device)
    kind, sc = b.get("gemm_kind", "tf32"), b.get("f16_scales", (1.0, 1.0, 1.0, 1.0))
    if fused:  # operands already in tail_s / tail_v (stages: 4 activated + 4 input scalar stages, gated + input vectors)
        base, a_v, comp = topo.tail_s.data_ptr(), topo.tail_v.data_ptr(), 2 * rows_all * 32
        st_s, st_v = 8, 2
    else:
        base = topo.a_ws.data_ptr()  # the conv operand is dead once the contraction has run
        a_v = base + 4 * st_s * rows_all * 32
        comp = st_v * rows_all * 32
    if fused:
        pass
    elif deferred:
        ops.tail_pack(topo.conv, None, x_in, b["s_in"], b["v_in"], b["c_act"], b["c_gate"], rows_all, base, a_v, comp,
                      rowptr=topo.rowptr, rhat=topo.rhat, t_edge=topo.t_edge, p2_scale=b["alpha1"], conv_has_v=bool(b["v_in"]))
    else:
        ops.tail_pack(topo.conv, vadd, x_in, b["s_in"], b["v_in"], b["c_act"], b["c_gate"], rows_all, base, a_v, comp)
    # (few row tiles -- small batches -- keep tail_mix: it spreads over all SMs, the GEMM's epilogue only over its few CTAs;
    # measured on C1, 11 tiles: 1.73 M atom-steps/s unfused vs 1.57 M fused)
    if kind == "f16" and os.environ.get("JAMUN_B200_TAIL_FUSE", "1") == "1" and (fused or rows_all // 128 > 74):
        # skip-mix + next-block scaling + operand packing in the GEMM's epilogue (mode 2): no tail_mix launch, no y round trip
        _ensure_tail_operands(topo)
        pack = x_scaled is not None
        epi = dict(mode=2, op_s=topo.tail_s.data_ptr() if pack else None, op_v=topo.tail_v.data_ptr() if pack else None,
                   op_v_comp_stride=2 * rows_all * 32, op_rows_pad=rows_all, x_res=None if skip_w is None else x_res.data_ptr(),
                   skip_w=None if skip_w is None else skip_w.data_ptr(), s_next=None if s_next is None else s_next.data_ptr(),
                   x_new=x_new.data_ptr(), x_scaled=x_scaled.data_ptr() if pack else None)
        ops.gemm_f16x3_fused([base] + [a_v + 4 * c * comp for c in range(3)],
                             [b["tail_bs_img"].data_ptr()] + [b["tail_bv_img"].data_ptr()] * 3, [st_s, st_v, st_v, st_v],
                             [128, 32, 32, 32], [120, 32, 32, 32], [0, 120, 152, 184], [1.0 / sc[2]] + [1.0 / sc[3]] * 3, N, rows_all, None,
                             epi, status=topo.gemm_status)
        topo.xs_op_of = x_scaled if pack else None
        return
    _gemm(topo, kind, [base] + [a_v + 4 * c * comp for c in range(3)],
          [b["tail_bs_img"].data_ptr()] + [b["tail_bv_img"].data_ptr()] * 3, [st_s, st_v, st_v, st_v], [128, 32, 32, 32],
          [120, 32, 32, 32], [0, 120, 152, 184], [1.0 / sc[2]] + [1.0 / sc[3]] * 3, N, rows_all, None, topo.ytail.data_ptr(), ops.HID)
    pack = x_scaled is not None and getattr(topo, "xs_op", None) is not None
    ops.tail_mix(topo.ytail, x_res, skip_w, s_next, x_new, x_scaled, topo.xs_op if pack else None, rows_all)
    topo.xs_op_of = x_scaled if pack else None


def e3conv_forward(plan: E3ConvPlan, topo: Topology, p: torch.Tensor, r_cut: float, g_out: torch.Tensor,
                   mu: Optional[torch.Tensor] = None, step: Optional[float] = None) -> torch.Tensor:
    """One evaluation of the network on scaled positions p over topo's current CSR -> g_out [N,3]."""
    if mu is None:
        mu, step = plan.radial_grid(r_cut)
    ops.edge_geom(p, topo.rowptr, topo.col, topo.edst, mu, step, topo.rhat, topo.rb)
    key = (plan.serial,)
    if topo.x0_key != key:  # constant per (topology, plan): embedding x initial noise scaling
        idx = list(topo.idx)
        if not plan.use_residue_sequence_index:
            idx[3] = None
        ops.atom_embed(idx, plan.tables, plan.s_init, topo.x0)
        topo.x0_key = key
    x_in, x_res = topo.x0, None
    nb = len(plan.blocks)
    # radial hidden of every layer in one pass over the edges (h_all: [L, cap, 64])
    if getattr(topo, "h_all", None) is None or topo.h_all.shape[0] != nb:
        topo.h_all = torch.zeros(nb, topo.cap, ops.EDGE_HID, dtype=torch.float32, device=topo.device)
    if plan.w0r_frag is not None:
        ops.edge_radial_hidden_mma(topo.rb, topo.ebond, topo.rowptr, plan.w0r_frag, plan.b0eff_all, topo.h_all)
    else:
        ops.edge_radial_hidden_all(topo.rb, topo.ebond, topo.rowptr, plan.w0r_all, plan.b0eff_all, topo.h_all)
    for l, b in enumerate(plan.blocks):
        vadd = None
        topo.h = topo.h_all[l]
        if CONV_IMPL == "simt":
            ops.conv_fwd(x_in, b["s_in"], b["v_in"], topo.rowptr, topo.col, topo.h, topo.rhat, b["m0"], b["m1"],
                         b["alpha0"], b["alpha1"], topo.conv)
        else:
            tail_tc = os.environ.get("JAMUN_B200_TAIL", TAIL_IMPL) == "tc" and os.environ.get("JAMUN_B200_BUILD", BUILD_IMPL) == "tc"
            conv_tc(topo, b, x_in, topo.conv, y_const_key=key if l == 0 else None, defer_reduce=tail_tc,
                    fuse_gate=tail_tc and l > 0 and topo.xs_op_of is x_in  # x_in's packed halves were written by the last tail
                    and _fuse_gate_ok(topo, b, os.environ.get("JAMUN_B200_BUILD", BUILD_IMPL)))
            vadd = conv_tc_join(topo, b)
        x_new, x_scaled = topo.xa[l & 1], topo.xs[l & 1]
        skip_w = plan.skips[l - 1] if l > 0 else None
        s_next = plan.scales[l] if l < nb - 1 else None
        block_tail(topo, b, x_in, x_res, skip_w, s_next, x_new, x_scaled if l < nb - 1 else None, vadd)
        x_in, x_res = x_scaled, x_new
    ops.head(x_res, plan.head_w1s, plan.head_w1v, plan.head_w2, plan.head_cgate, g_out)
    return g_out
