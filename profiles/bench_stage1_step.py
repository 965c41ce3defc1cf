import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from kidney_diffusion_b200 import ops
from kidney_diffusion_b200.factories import init_imagen_ultra_res, randomize_zero_init_
from kidney_diffusion_b200.imagen import CounterNoise
dev = torch.device("cuda:0")
torch.manual_seed(0)
im = init_imagen_ultra_res(1, 1, version="v_param"); randomize_zero_init_(im); im = im.to(dev).eval()
noise = CounterNoise(1, 0)
for B in (1, 2, 10, 16):
    S = 64
    cond = torch.rand(B, 3, 1024, 1024, device=dev)
    for inp in (False, True):
        kw = {}
        if inp:
            kw = dict(inpaint_images=torch.rand(B, 3, S, S, device=dev), inpaint_masks=(torch.rand(B, S, S, device=dev) > 0.5), inpaint_resample_times=1)
        run = im.stage_run(1, (B, 3, S, S), noise=noise, cond_images=cond, **kw)
        for k in range(3): run.step(k)
        torch.cuda.synchronize(); t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(3, 23): run.step(k)
        e1.record(); torch.cuda.synchronize()
        print(f"U1 B={B:2d} inpaint={inp}: {e0.elapsed_time(e1)/20:.2f} ms/step (wall {(time.time()-t0)/20*1e3:.2f})")
