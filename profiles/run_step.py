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
U = int(os.environ.get("KD_PROFILE_UNET", "3"))   # cascade stage: 1 (64^2 base), 2 (256^2), 3 (1024^2)
S = int(os.environ.get("KD_PROFILE_SIZE", str({1: 64, 2: 256, 3: 1024}[U])))
dev = torch.device("cuda:0")
torch.manual_seed(0)
imagen = init_imagen_ultra_res(1, U, version="v_param")
randomize_zero_init_(imagen)
imagen = imagen.to(dev).eval()
imagen.use_cuda_graph = False
noise = CounterNoise(1234, 0)
g = torch.Generator().manual_seed(1)
cond = torch.rand(B, 3, S, S, generator=g).to(dev)
lowres = torch.randn(B, 3, S, S, generator=g).to(dev)
run = imagen.stage_run(U, (B, 3, S, S), noise=noise, lowres_cond_img=lowres if U > 1 else None,
                       lowres_noise_level=0.2 if U > 1 else None, cond_images=cond)
run.step(0)
run.step(1)
torch.cuda.synchronize()
n0 = ops.launch_count
torch.cuda.profiler.start()
run.step(2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in profiled step:", ops.launch_count - n0)

if os.environ.get("KD_CONV_IMPL"):
    ops.set_conv_impl(int(os.environ["KD_CONV_IMPL"]))
if os.environ.get("KD_CONV_TABLE"):
    import collections

    ops.conv_profile = []
    run.step(3)
    torch.cuda.synchronize()
    prof, ops.conv_profile = ops.conv_profile, None
    agg = collections.OrderedDict()
    for flops, e0, e1, label in prof:
        a = agg.setdefault(label, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += flops
        a[2] += e0.elapsed_time(e1)
    print("mode B H W Cin Cout k | launches | GFLOP each | ms each | TFLOP/s")
    for label, (n, fl, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print(f"{label} | {n:3d} | {fl / n / 1e9:8.1f} | {ms / n:7.3f} | {fl / ms / 1e9:7.1f}")
    tot_f, tot_ms = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())
    print(f"total conv: {tot_f / 1e12:.3f} TFLOP in {tot_ms:.2f} ms = {tot_f / tot_ms / 1e9:.1f} TFLOP/s")

if os.environ.get("KD_OP_TABLE"):
    import collections

    ops.op_profile = []
    run.step(4)
    torch.cuda.synchronize()
    prof, ops.op_profile = ops.op_profile, None
    agg = collections.OrderedDict()
    for name, e0, e1, nbytes in prof:
        a = agg.setdefault(name, [0, 0.0, 0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
        a[2] += nbytes
    tot = sum(a[1] for a in agg.values())
    print(f"op table, B={B}, S={S}: total {tot:.2f} ms ({tot / B:.2f} ms per patch-step)")
    for name, (n, ms, nbytes) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {name:16s} n={n:4d}  {ms:8.3f} ms  {100 * ms / tot:5.1f}%   {nbytes / 1e9:7.2f} GB touched  {nbytes / ms / 1e6:6.0f} GB/s")
