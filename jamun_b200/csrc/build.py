"""Build the jamun_b200 CUDA library in-tree: nvcc -> jamun_b200/csrc/libjamun_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SOURCES = ["graph.cu", "conv_simt.cu", "langevin.cu", "conv_build.cu", "conv_build_tc.cu", "gemm_tf32x3.cu", "pack.cu", "tensor_product.cu", "train_generic.cu", "train_conv_bwd.cu", "train_misc.cu", "gemm_atb.cu"]
LIB = HERE / "libjamun_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [HERE / s for s in SOURCES] + list(HERE.glob("*.cuh")) + [HERE.parent.parent / "include" / "jamun_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objs = []
    for src in SOURCES:
        obj = HERE / (src[:-3] + ".o")
        cmd = [NVCC, *FLAGS, "-c", str(HERE / src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        (HERE / (src[:-3] + ".ptxas.log")).write_text(res.stderr)
        objs.append(str(obj))
    cmd = [NVCC, "-shared", "-o", str(LIB), *objs, "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
