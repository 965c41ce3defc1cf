"""Build libkidney_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The shared library is a plain C ABI (include/kidney_b200.h); it links only the CUDA runtime (statically) and looks up
cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, so it has no link-time dependency on libcuda or torch.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkidney_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")

SOURCES = ["kd_abi.cu", "kd_conv_gemm.cu", "kd_norm_gca.cu", "kd_cond_attn.cu", "kd_sampler.cu", "kd_edge_convs.cu", "kd_init_conv.cu", "kd_experiments.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _source_hash() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, "..", "include", "kidney_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str | None:
    cand = shutil.which("nvcc")
    if cand:
        return cand
    for p in ("/usr/local/cuda/bin/nvcc",):
        if os.path.exists(p):
            return p
    return None


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into libkidney_b200.so; no-op when sources are unchanged."""
    want = _source_hash()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == want:
        return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libkidney_b200.so")
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libkidney_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(want)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
