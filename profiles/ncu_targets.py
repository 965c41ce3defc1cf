"""One shipped kernel variant per invocation, launched a few times on the shape it has inside the B = 8 1024^2 patch-step, for
`ncu --set full -k regex:<kernel> -s 2 -c 1` (profiles/README.md).  Also prints the CUDA-event time and algorithmic TFLOP/s / GB/s.
Targets: pre128 pre256 res1x1 c1x1 shuffle init final gca_pool attn cublas
Usage: python profiles/ncu_targets.py <target> [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kidney_diffusion_b200 import ops
from kidney_diffusion_b200.build import build_library

build_library()
target = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = "cuda"


def act(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).half()


def run(fn, flops=0.0, nbytes=0.0, n=4):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{target} B={B}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TFLOP/s  {nbytes / ms / 1e6:.0f} GB/s (algorithmic)", flush=True)


def pre(S, C):
    xa, w, bias = act(B, S, S, C), act(C, 9 * C), torch.randn(C, device=dev)
    st = ops.oct_stats(xa)
    gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
    _, coef = ops.gn_finalize_oct(st, 1.0, None, 1.0, C // 8, 8, count=(C // 8) * S * S, gamma=gamma, beta=beta, want_coef=True)
    px = B * S * S
    run(lambda: ops.conv_gemm(xa, w, bias, ksize=3, want_stats=True, pre_coef=coef), 2.0 * px * 9 * C * C, px * 2 * 2 * C)


if target == "pre128":      # conv_gemm_halo_kernel<128, 4, 4, 0, 3, 1>: 3x3 128 -> 128 at 1024^2, fused GroupNorm + SiLU input
    pre(1024, 128)
elif target == "pre256":    # conv_gemm_halo_kernel<256, 3, 7, 0, 1, 1>: 3x3 256 -> 256 at 256^2
    pre(256, 256)
elif target == "res1x1":    # conv_gemm_pair_kernel<128, 6, 1>: residual 1x1 over the skip concat + gate * h2 addend at 1024^2 (HBM-bound)
    S = 1024
    xa, xb, w, bias, h2 = act(B, S, S, 128), act(B, S, S, 128), act(128, 256), torch.randn(128, device=dev), act(B, S, S, 128)
    gate = torch.rand(B, 128, device=dev)
    px = B * S * S
    run(lambda: ops.conv_gemm(xa, w, bias, xb=xb, ksize=1, addend=h2, addend_scale=gate, want_stats=True), 2.0 * px * 256 * 128, px * 2 * 512)
elif target == "c1x1":      # conv_gemm_pair_kernel<256, 5, 0>: 1x1 2048 -> 1024 at 64^2 (tensor-bound 1x1)
    S = 64
    xa, w, bias = act(B, S, S, 2048), act(1024, 2048), torch.randn(1024, device=dev)
    px = B * S * S
    run(lambda: ops.conv_gemm(xa, w, bias, ksize=1, want_stats=True), 2.0 * px * 2048 * 1024, px * 2 * 3072)
elif target == "init":      # init_conv_kernel<128>: CrossEmbed 15x15 merged filter, 3 image channels, 1024^2
    S = 1024
    x = torch.randn(B, 3, S, S, device=dev)
    wd = act(128, ops.init_conv_kp(3, 15))
    out = torch.empty(B, S, S, 128, device=dev, dtype=torch.float16)
    taps = (9 * 64 + 49 * 32 + 225 * 32) / 128
    run(lambda: ops.init_conv(x, 15, wd, torch.randn(128, device=dev), None, out, algo_taps=taps), 2.0 * B * S * S * 128 * 3 * taps,
        B * S * S * (12 + 256))
elif target == "final":     # final_conv_kernel: 3x3, 128 + 3 -> 3 channels at 1024^2 (HBM-bound)
    S = 1024
    xa, xb = act(B, S, S, 128), torch.randn(B, 3, S, S, device=dev)
    w, bias = torch.randn(3, 3, 3, 131, device=dev) * 0.03, torch.randn(3, device=dev)
    run(lambda: ops.final_conv(xa, xb, w, bias), 2.0 * B * S * S * 9 * 131 * 3, B * S * S * (256 + 12 + 12))
elif target == "gca_pool":  # gca_pool_kernel on the 512^2 x 128 tensor (HBM-bound)
    S = 512
    x = act(B, S, S, 128)
    logits = torch.randn(2, B, S * S, device=dev)
    run(lambda: ops.gca_pool(x, logits), 0.0, B * S * S * (256 + 8))
elif target == "attn":      # attn_mqa_tc_kernel (tcgen05; set KD_ATTN_LEGACY=1 for the mma.sync kernel): N = 4096 tokens, 8 heads of 64, multi-query
    if os.environ.get("KD_ATTN_LEGACY"):
        ops.ATTN_TC_MIN_TOKENS = 1 << 30
    N = 4096
    qkv = act(B, N, 512 + 128)
    kv = ops.kv_assemble(qkv, 512, None, torch.randn(2, 64, device=dev))
    run(lambda: ops.attn_mqa(qkv, kv, 8, 0.125), 4.0 * B * 8 * N * (N + 1) * 64, B * N * 2 * (640 + 512))
elif target == "shuffle":   # conv_gemm_pair_kernel<256, 5, 0>: PixelShuffleUpsample 1x1 128 -> 512 + SiLU at 512^2, stored as 1024^2 x 128
    S = 512
    xa, w, bias = act(B, S, S, 128), act(512, 128), torch.randn(512, device=dev)
    px = B * S * S
    run(lambda: ops.conv_gemm(xa, w, bias, ksize=1, act=ops.ACT_SILU, out_mode=1), 2.0 * px * 128 * 512, px * 2 * (128 + 512))
elif target == "cublas":    # calibration of the tensor-pipe counter: the library GEMM MEASURED_PEAKS.json's peak comes from
    a, b = torch.randn(8192, 8192, device=dev).bfloat16(), torch.randn(8192, 8192, device=dev).bfloat16()
    run(lambda: a @ b, 2.0 * 8192 ** 3, 3 * 8192 * 8192 * 2)
else:
    raise SystemExit(f"unknown target {target}")
