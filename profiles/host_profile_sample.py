"""Host-side profile of Imagen.sample() for one 1024^2 patch with a 5-step schedule (the per-batch call of the patch-grid executor):
where does the host spend its time, and does it run ahead of the GPU?  Usage: python profiles/host_profile_sample.py"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kidney_diffusion_b200.factories import init_imagen_ultra_res, randomize_zero_init_

dev = torch.device("cuda:0")
torch.manual_seed(0)
im = init_imagen_ultra_res(1, 3, version="v_param", timesteps=(8, 4, 5))
randomize_zero_init_(im)
im = im.to(dev).eval()
B, S = 1, 1024
cond = torch.rand(B, 3, S, S, device=dev)
start = torch.rand(B, 3, 256, 256, device=dev)
inp = torch.rand(B, 3, S, S, device=dev)
mask = torch.zeros(B, S, S, device=dev)
mask[:, :256] = 1
kw = dict(batch_size=B, cond_images=cond, start_image_or_video=start, start_at_unet_number=3, stop_at_unet_number=3, inpaint_images=inp,
          inpaint_masks=mask, inpaint_resample_times=1, use_tqdm=False, device=dev, return_pil_images=False)
for i in range(3):
    im.sample(noise_key=[i], **kw)
torch.cuda.synchronize()
N = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
pr = cProfile.Profile()
t0 = time.time()
e0.record()
pr.enable()
for i in range(N):
    im.sample(noise_key=[10 + i], **kw)
pr.disable()
t_issue = time.time() - t0
e1.record()
torch.cuda.synchronize()
print(f"{N} sample() calls of 5 steps, B = {B}: host issue time {t_issue * 1e3 / N:.1f} ms per call, GPU time {e0.elapsed_time(e1) / N:.1f} ms per call")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
