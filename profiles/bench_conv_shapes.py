"""CUDA-event timing of single kd_conv_gemm shapes of the 1024^2 patch-step (B patches), with algorithmic FLOPs and HBM bytes.
Usage: python profiles/bench_conv_shapes.py [B] [case ...]   (cases: res1x1 shuffle c3_128 c3_256 all)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kidney_diffusion_b200 import ops
from kidney_diffusion_b200.build import build_library

build_library()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cases = sys.argv[2:] or ["all"]
if os.environ.get("KD_CONV_IMPL"):
    ops.set_conv_impl(int(os.environ["KD_CONV_IMPL"]))
dev = "cuda"


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def act(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).half()


def report(name, ms, flops, nbytes):
    print(f"{name:44s} {ms:7.3f} ms  {flops / ms / 1e9:7.0f} TFLOP/s  {nbytes / ms / 1e6:6.0f} GB/s", flush=True)


def want(c):
    return "all" in cases or c in cases


if want("res1x1"):
    S = 1024
    xa, xb = act(B, S, S, 128), act(B, S, S, 128)
    w = act(128, 256)
    bias = torch.randn(128, device=dev)
    h2 = act(B, S, S, 128)
    gate = torch.rand(B, 128, device=dev)
    px = B * S * S
    ms = timeit(lambda: ops.conv_gemm(xa, w, bias, xb=xb, ksize=1, addend=h2, addend_scale=gate, want_stats=True))
    report("1x1 128+128->128 @1024 +addend*gate +stats", ms, 2.0 * px * 256 * 128, px * 2 * (256 + 128 + 128))
    ms = timeit(lambda: ops.conv_gemm(xa, w, bias, xb=xb, ksize=1))
    report("1x1 128+128->128 @1024 plain", ms, 2.0 * px * 256 * 128, px * 2 * (256 + 128))
if want("shuffle"):
    S = 512
    xa = act(B, S, S, 128)
    w = act(512, 128)
    bias = torch.randn(512, device=dev)
    px = B * S * S
    ms = timeit(lambda: ops.conv_gemm(xa, w, bias, ksize=1, act=ops.ACT_SILU, out_mode=1))
    report("1x1 128->512 @512 SiLU + pixel shuffle", ms, 2.0 * px * 128 * 512, px * 2 * (128 + 512))
if want("c3_128"):
    S = 1024
    xa = act(B, S, S, 128)
    w = act(128, 9 * 128)
    bias = torch.randn(128, device=dev)
    px = B * S * S
    ms = timeit(lambda: ops.conv_gemm(xa, w, bias, ksize=3, want_stats=True))
    report("3x3 128->128 @1024 +stats", ms, 2.0 * px * 9 * 128 * 128, px * 2 * 256)
    add = act(B, S, S, 128)
    ms = timeit(lambda: ops.conv_gemm(xa, w, bias, ksize=3, addend=add, want_stats=True))
    report("3x3 128->128 @1024 +addend +stats", ms, 2.0 * px * 9 * 128 * 128, px * 2 * 384)
if want("c3_256"):
    S = 256
    xa = act(B, S, S, 256)
    w = act(256, 9 * 256)
    bias = torch.randn(256, device=dev)
    px = B * S * S
    ms = timeit(lambda: ops.conv_gemm(xa, w, bias, ksize=3, want_stats=True))
    report("3x3 256->256 @256 +stats", ms, 2.0 * px * 9 * 256 * 256, px * 2 * 512)
if want("pre"):
    for (S, Cin, Cout) in ((256, 256, 256), (1024, 128, 128)):
        xa = act(B, S, S, Cin)
        w = act(Cout, 9 * Cin)
        bias = torch.randn(Cout, device=dev)
        px = B * S * S
        st = ops.oct_stats(xa)
        gamma, beta = torch.randn(Cin, device=dev), torch.randn(Cin, device=dev)
        _, coef = ops.gn_finalize_oct(st, 1.0, None, 1.0, Cin // 8, 8, count=(Cin // 8) * S * S, gamma=gamma, beta=beta, want_coef=True)
        ms = timeit(lambda: ops.conv_gemm(xa, w, bias, ksize=3, want_stats=True, pre_coef=coef))
        report(f"3x3 {Cin}->{Cout} @{S} +stats, fused GroupNorm+SiLU input", ms, 2.0 * px * 9 * Cin * Cout, px * 2 * (Cin + Cout))
        add = act(B, S, S, Cout)
        ms = timeit(lambda: ops.conv_gemm(xa, w, bias, ksize=3, want_stats=True, pre_coef=coef, addend=add))
        report(f"3x3 {Cin}->{Cout} @{S} +stats, fused GN input + residual", ms, 2.0 * px * 9 * Cin * Cout, px * 2 * (Cin + 2 * Cout))
        ms = timeit(lambda: ops.conv_gemm(xa, w, bias, ksize=3, want_stats=True))
        report(f"3x3 {Cin}->{Cout} @{S} +stats, plain input", ms, 2.0 * px * 9 * Cin * Cout, px * 2 * (Cin + Cout))
