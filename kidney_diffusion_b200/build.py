"""Build libkidney_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The shared library is a plain C ABI (include/kidney_b200.h); it links only the CUDA runtime (statically) and looks up
cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, so it has no link-time dependency on libcuda or torch.
Every translation unit is compiled separately (in parallel, cached by a hash of its text + the shared headers) and linked.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libkidney_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")

SOURCES = ["kd_abi.cu", "kd_conv_gemm.cu", "kd_norm_gca.cu", "kd_cond_attn.cu", "kd_sampler.cu", "kd_edge_convs.cu", "kd_init_conv.cu",
           "kd_grid_peer.cu", "kd_guard.cu", "kd_linattn.cu", "kd_attn_tc.cu", "kd_precise.cu"]
CC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]
NVCC_FLAGS = CC_FLAGS + LINK_FLAGS  # (kept for the stamp)


def _headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + [
        os.path.join(HERE, "..", "include", "kidney_b200.h")]


def _hash(paths, extra="") -> str:
    h = hashlib.sha256()
    for f in paths:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(extra.encode())
    return h.hexdigest()


def _source_hash() -> str:
    files = [os.path.join(CSRC, s) for s in SOURCES] + _headers()
    return _hash(files, " ".join(NVCC_FLAGS))


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
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers()
    sources = list(SOURCES)

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        tag = _hash([path] + hdr, " ".join(CC_FLAGS))
        tagf = obj + ".hash"
        if not force and os.path.exists(obj) and os.path.exists(tagf) and open(tagf).read() == tag:
            return obj, ""
        cmd = [nvcc, *CC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", obj, path]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed compiling {src}")
        with open(tagf, "w") as fh:
            fh.write(tag)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as pool:
        results = list(pool.map(compile_one, sources))
    if verbose:
        sys.stderr.write("".join(r[1] for r in results))
    res = subprocess.run([nvcc, *LINK_FLAGS, "-o", LIB, *[r[0] for r in results]], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libkidney_b200.so")
    with open(STAMP, "w") as fh:
        fh.write(want)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
