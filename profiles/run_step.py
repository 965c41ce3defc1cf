"""Profiling driver: one eager (non-graph) 1024^2 patch-step of the config-3/4 SR UNet, B=1, bracketed by
cudaProfilerStart/Stop so `ncu --profile-from-start off` sees exactly one step (see profiles/README.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from kidney_diffusion_b200 import ops, schedule
from kidney_diffusion_b200.factories import init_imagen_ultra_res, randomize_zero_init_
from kidney_diffusion_b200.imagen import CounterNoise

B = int(os.environ.get("KD_PROFILE_BATCH", "1"))
S = int(os.environ.get("KD_PROFILE_SIZE", "1024"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
imagen = init_imagen_ultra_res(1, 3, version="v_param")
randomize_zero_init_(imagen)
imagen = imagen.to(dev).eval()
imagen.use_cuda_graph = False
noise = CounterNoise(1234, 0)
g = torch.Generator().manual_seed(1)
cond = torch.rand(B, 3, S, S, generator=g).to(dev)
lowres = torch.randn(B, 3, S, S, generator=g).to(dev)
run = imagen.stage_run(3, (B, 3, S, S), noise=noise, lowres_cond_img=lowres, lowres_noise_level=0.2, cond_images=cond)
run.step(0)
run.step(1)
torch.cuda.synchronize()
n0 = ops.launch_count
torch.cuda.profiler.start()
run.step(2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in profiled step:", ops.launch_count - n0)
