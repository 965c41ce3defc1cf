"""Micro-benchmark of the non-GEMM-shaped kernels of a patch-step at full size (B patches of 1024^2): CUDA-event timing of
kd_init_conv, kd_final_conv and the global-context pair, with their algorithmic HBM bytes.  Usage: python profiles/bench_edge.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kidney_diffusion_b200 import ops
from kidney_diffusion_b200.build import build_library

build_library()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S, dim, dev = 1024, 128, "cuda"


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


x = torch.randn(B, 3, S, S, device=dev)
wp = (torch.randn(dim, ops.init_conv_kp(3, 15), device=dev) * 0.05).half()
bias = torch.randn(dim, device=dev)
add = torch.randn(B, S, S, dim, device=dev).half()
out = torch.empty_like(add)
ms = timeit(lambda: ops.init_conv(x, 15, wp, bias, add, out))
gb = (x.numel() * 4 + add.numel() * 2 + out.numel() * 2) / 1e9
print(f"init_conv  B={B}: {ms:.3f} ms  ({gb / ms * 1e3:.0f} GB/s algorithmic, {2 * B * S * S * dim * 768 / ms / 1e9:.0f} TFLOP/s issued)")
ms = timeit(lambda: ops.init_conv(x, 15, wp, bias, None, out))
print(f"init_conv (no addend): {ms:.3f} ms")

w = torch.randn(3, 3, 3, dim + 3, device=dev) * 0.05
fb = torch.randn(3, device=dev)
ms = timeit(lambda: ops.final_conv(add, x, w, fb))
gb = (add.numel() * 2 + 2 * x.numel() * 4) / 1e9
print(f"final_conv B={B}: {ms:.3f} ms  ({gb / ms * 1e3:.0f} GB/s algorithmic)")

# GlobalContext pooling: kd_gca_pool + kd_gca_finalize per level of the SR UNet
import ctypes
from kidney_diffusion_b200.ops import _ptr, _stream, lib
for (S2, C) in ((512, 128), (256, 256), (128, 512), (64, 1024)):
    xg = torch.randn(B, S2, S2, C, device=dev).half()
    lg = torch.randn(max(1, C // 64), B, S2 * S2, device=dev)
    HW = S2 * S2
    nblk = lib().kd_elementwise_blocks(HW, C)
    part = torch.empty((B, nblk, C), device=dev)
    ml = torch.empty((B, nblk, 2), device=dev)
    pooled = torch.empty((B, C), device=dev)
    t_pool = timeit(lambda: lib().kd_gca_pool(_ptr(xg), _ptr(lg), lg.shape[0], B, HW, C, nblk, _ptr(part), _ptr(ml), _stream()), 10)
    t_fin = timeit(lambda: lib().kd_gca_finalize(_ptr(part), _ptr(ml), B, nblk, C, _ptr(pooled), _stream()), 10)
    gb = xg.numel() * 2 / 1e9
    print(f"gca_pool {S2}^2 x {C}: pool {t_pool * 1e3:7.1f} us ({gb / t_pool * 1e3:5.0f} GB/s, nblk {nblk}), finalize {t_fin * 1e3:6.1f} us")
