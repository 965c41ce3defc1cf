"""In-graph time of one patch-step (UNet CUDA graph + dynamic threshold + update, with inpainting as in the grid sampler) for
chosen (stage, batch) points.  Usage: python profiles/step_time.py 1:1,2,4,16 2:1,8 3:1,2"""
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
noise = CounterNoise(1, 0)
out = {}
for spec in sys.argv[1:]:
    U, bs = spec.split(":")
    U = int(U)
    torch.manual_seed(0)
    im = init_imagen_ultra_res(1, U, version="v_param", timesteps=(16, 16, 16))
    randomize_zero_init_(im)
    im = im.to(dev).eval()
    S = {1: 64, 2: 256, 3: 1024}[U]
    out[U] = {}
    for B in (int(b) for b in bs.split(",")):
        cond = torch.rand(B, 3, 1024, 1024, device=dev)
        lowres = torch.randn(B, 3, S, S, device=dev) if U > 1 else None
        mask = torch.zeros(B, S, S, device=dev)
        mask[:, : S // 4] = 1
        run = im.stage_run(U, (B, 3, S, S), noise=noise, lowres_cond_img=lowres, lowres_noise_level=0.2 if U > 1 else None, cond_images=cond,
                           inpaint_images=torch.rand(B, 3, S, S, device=dev), inpaint_masks=mask, inpaint_resample_times=1)
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
        out[U][B] = round(e0.elapsed_time(e1) / n, 3)
        print(f"stage {U} B={B}: {out[U][B]} ms per step ({out[U][B] / B:.3f} per patch-step)", flush=True)
        del run, cond, lowres
        im._graphs.clear()
        torch.cuda.empty_cache()
    del im
    torch.cuda.empty_cache()
print(json.dumps(out))
