"""Device-side cost of ONE Imagen.sample() call of the 1024^2 stage outside its sampling steps (the per-batch overhead of the
patch-grid executor): time sample() with 2 / 5 / 10 steps (slope = step, intercept = per-call work) and list the kernels of one
call by total time (torch.profiler, CUDA activities).  Usage: python profiles/sample_call_profile.py"""
import os
import sys

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
res = {}
for steps in (2, 5, 10):
    im.noise_schedulers[2].num_timesteps = steps
    for i in range(3):
        im.sample(noise_key=[i], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8):
        im.sample(noise_key=[10 + i], **kw)
    e1.record()
    torch.cuda.synchronize()
    res[steps] = e0.elapsed_time(e1) / 8
slope = (res[10] - res[2]) / 8
print(f"sample() GPU time per call: {res}; step = {slope:.2f} ms, per-call work outside the steps = {res[5] - 5 * slope:.2f} ms")
im.noise_schedulers[2].num_timesteps = 2
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    im.sample(noise_key=[99], **kw)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof.key_averages()]
rows = [r for r in rows if r[1] > 0]
rows.sort(key=lambda r: -r[1])
print("device time by kernel / op for one 2-step call (us, count):")
for k, t, c in rows[:32]:
    print(f"  {t:10.1f}  n={c:3d}  {k[:110]}")
