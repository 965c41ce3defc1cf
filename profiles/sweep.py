"""BASELINE configs[4]: UNet denoise-step sweep, batch 1-64 x resolution 64 / 256 / 1024 (the three cascade stages of the
ultra-res models), one JSON line per point: CUDA-graph replay time of a whole patch-step (UNet + exact dynamic threshold +
update), algorithmic TFLOP/s and the fraction of the measured sustained tensor peak.
Usage: python profiles/sweep.py > profiles/r01_sweep.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kidney_diffusion_b200.build import build_library
from kidney_diffusion_b200.factories import init_imagen_ultra_res, randomize_zero_init_
from kidney_diffusion_b200.imagen import CounterNoise

build_library()
dev = torch.device("cuda:0")
peak = 1388.6
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        mp = json.load(f)
    peak = float(mp.get("bf16_tflops_sustained", mp.get("tf_sustained", peak)))
except Exception:
    pass
SIZES = {1: 64, 2: 256, 3: 1024}
MAXB = {1: 64, 2: 64, 3: 32}  # 1024^2: 3.3 GB of graph-private activations per patch
noise = CounterNoise(1, 0)
for U in (1, 2, 3):
    torch.manual_seed(0)
    im = init_imagen_ultra_res(1, U, version="v_param")
    randomize_zero_init_(im)
    im = im.to(dev).eval()
    S = SIZES[U]
    flops = None
    for B in (1, 2, 4, 8, 16, 32, 64):
        if B > MAXB[U]:
            continue
        cond = torch.rand(B, 3, 1024, 1024, device=dev) if U < 3 else torch.rand(B, 3, S, S, device=dev)
        lowres = torch.randn(B, 3, S, S, device=dev) if U > 1 else None
        run = im.stage_run(U, (B, 3, S, S), noise=noise, lowres_cond_img=lowres, lowres_noise_level=0.2 if U > 1 else None, cond_images=cond)
        for k in range(3):
            run.step(k)
        torch.cuda.synchronize()
        n = 10 if U < 3 else 4
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(3, 3 + n):
            run.step(k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if flops is None:  # algorithmic conv / linear FLOPs of one sample, counted once per stage by the op layer
            from kidney_diffusion_b200 import ops
            ops.conv_profile = []
            im.use_cuda_graph = False
            run.step(3 + n)
            torch.cuda.synchronize()
            flops = sum(p[0] for p in ops.conv_profile) / B
            ops.conv_profile = None
            im.use_cuda_graph = True
        tf = B * flops / (ms / 1e3) / 1e12
        print(json.dumps(dict(stage=U, resolution=S, batch=B, ms_per_step=round(ms, 3), ms_per_patch_step=round(ms / B, 4),
                              patch_steps_per_s=round(B / ms * 1e3, 2), gflop_per_patch_step=round(flops / 1e9, 1),
                              tensor_tflops=round(tf, 1), frac_of_sustained_peak=round(tf / peak, 3))), flush=True)
        del run, cond, lowres
        im._graphs.clear()
        torch.cuda.empty_cache()
    del im
    torch.cuda.empty_cache()
