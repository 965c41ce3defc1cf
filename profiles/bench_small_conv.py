"""Fixed cost of a tensor-core conv launch on tiny problems (the 64^2 base stage is made of these): back-to-back launches under a
CUDA graph, alone and interleaved with a small elementwise kernel.  Usage: python profiles/bench_small_conv.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kidney_diffusion_b200 import ops
from kidney_diffusion_b200.build import build_library

build_library()
dev = "cuda"


def graph_time(fn, reps=50):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / reps * 1e3  # us per call


# programmatic dependent launch on / off: run with KD_NO_PDL=1 in the environment (read once by libkidney_b200 at load)
for (B, S, Cin, Cout, k) in ((2, 8, 512, 1024, 1), (2, 8, 1792, 1024, 1), (2, 8, 768, 1024, 3), (2, 16, 768, 768, 3), (2, 32, 256, 256, 3), (2, 64, 512, 256, 3),
                             (16, 8, 768, 1024, 3), (16, 16, 768, 768, 3)):
    x = (torch.randn(B, S, S, Cin, device=dev) * 0.5).half()
    w = (torch.randn(Cout, k * k * Cin, device=dev) * 0.05).half()
    b = torch.randn(Cout, device=dev)
    y = torch.empty(B, S, S, Cout, device=dev, dtype=torch.float16)
    t_conv = graph_time(lambda: ops.conv_gemm(x, w, b, ksize=k, out=y))
    t_mix = graph_time(lambda: (ops.conv_gemm(x, w, b, ksize=k, out=y), ops.gate_residual(y, None, None)))
    t_ew = graph_time(lambda: ops.gate_residual(y, None, None))
    gf = 2.0 * B * S * S * Cout * k * k * Cin / 1e9
    print(f"B={B:2d} {S:3d}^2 {Cin:4d}->{Cout:4d} k{k}: conv {t_conv:6.1f} us ({gf / t_conv * 1e-3:6.1f} TFLOP/s), conv+elementwise {t_mix:6.1f} us, elementwise {t_ew:5.1f} us")
