"""BASELINE.json configs[0] and configs[1] end to end on one B200 (configs[2] / [3] are bench.py, configs[4] is profiles/sweep.py).

cfg1  unconditional 64x64 base UNet (train_uncond.py:30-37 shape with dim=128), random init, 50-step DDPM sampling, batch 4:
      Imagen.sample on the GPU vs the SAME model (identical weights, identical injected noise) on the fp32 CPU oracle: final-sample
      rel-L2 and both wall times (the config "runs on CPU today").
cfg2  mask + clinical-vector conditioned cascade 64 -> 256 (train.py:28-52 full width, text_embeds [0.0, 0.5, 0.2], 4-channel one-hot
      label map at 1024^2, batch 3, full 1024 + 256 steps): GPU wall time and patch-steps/s per stage.
One JSON line per config.  Usage: python profiles/configs.py [--no-cpu]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from kidney_diffusion_b200 import Imagen, Unet
from kidney_diffusion_b200.build import build_library
from kidney_diffusion_b200.factories import cond_unet, randomize_zero_init_

build_library()
dev = torch.device("cuda:0")


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


def cfg1(no_cpu):
    from helpers import KeyedNoise, rel_l2
    from oracle import imagen_oracle as O

    kw = dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=3, layer_attns=(False, True, True, True), layer_cross_attns=(False, True, True, True))
    torch.manual_seed(0)
    oi = O.Imagen(unets=(O.Unet(**kw),), image_sizes=(64,), timesteps=50, condition_on_text=False).eval()
    O.randomize_zero_init_(oi)
    pi = Imagen(unets=(Unet(**kw),), image_sizes=(64,), timesteps=50, condition_on_text=False)
    pi.load_state_dict(oi.state_dict())
    pi = pi.to(dev).eval()
    kn = KeyedNoise(3)
    pi.noise_fn = kn.dev
    _, cold = timed(lambda: pi.sample(batch_size=4, use_tqdm=False, device=dev))
    out, warm = timed(lambda: pi.sample(batch_size=4, use_tqdm=False, device=dev))
    line = dict(config="cfg1: unconditional 64x64 base UNet dim=128, 50-step DDPM, batch 4, random init", gpu_seconds=warm, gpu_seconds_first_call=cold,
                gpu_patch_steps_per_s=4 * 50 / warm, params_M=sum(p.numel() for p in pi.parameters()) / 1e6, dtype="f16", finite=bool(torch.isfinite(out).all()))
    if not no_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        ref = oi.sample(batch_size=4, noise_fn=kn.cpu)
        cpu_s = time.perf_counter() - t0
        line.update(cpu_seconds=cpu_s, cpu_cores=os.cpu_count(), cpu_kind="port (fp32 PyTorch oracle)", speedup=cpu_s / warm,
                    final_sample_rel_l2_vs_oracle=rel_l2(out, ref), tolerance=1e-2)
        pi.set_precision("fp32")  # the precise CUDA-core path on the same model and noise
        timed(lambda: pi.sample(batch_size=4, use_tqdm=False, device=dev))
        out32, s32 = timed(lambda: pi.sample(batch_size=4, use_tqdm=False, device=dev))
        line.update(fp32_path_seconds=s32, fp32_path_final_sample_rel_l2_vs_oracle=rel_l2(out32, ref), fp32_tolerance=1e-4,
                    fp16_path_vs_fp32_path_rel_l2=rel_l2(out, out32))
        pi.set_precision("fp16")
    print(json.dumps(line), flush=True)


def cfg2():
    torch.manual_seed(1)
    B = 3
    pi = Imagen(unets=(cond_unet(1), cond_unet(2)), image_sizes=(64, 256), timesteps=(1024, 256), pred_objectives=("noise", "v"), text_embed_dim=3,
                random_crop_sizes=(None, None))
    randomize_zero_init_(pi)
    pi = pi.to(dev).eval()
    conds = torch.tensor([0.0, 0.5, 0.2]).reshape(1, 1, 3).repeat_interleave(B, dim=0).to(dev)
    labels = torch.randint(0, 5, (B, 1024, 1024), generator=torch.Generator().manual_seed(1))
    deep = torch.stack([(labels == k).float() for k in range(1, 5)], dim=1).to(dev)
    stage = {}
    for u, steps in ((1, 1024), (2, 256)):
        start = None if u == 1 else torch.rand(B, 3, 64, 64, device=dev)
        fn = lambda: pi.sample(text_embeds=conds, cond_images=deep, start_image_or_video=start, start_at_unet_number=u, stop_at_unet_number=u,
                               use_tqdm=False, device=dev, noise_key=7)
        timed(fn)
        out, s = timed(fn)
        stage[u] = dict(seconds=s, steps=steps, ms_per_step=1e3 * s / steps, patch_steps_per_s=B * steps / s, finite=bool(torch.isfinite(out).all()))
    full, total = timed(lambda: pi.sample(text_embeds=conds, cond_images=deep, use_tqdm=False, device=dev, noise_key=7))
    print(json.dumps(dict(config="cfg2: mask + clinical-vector conditioned cascade 64 -> 256 (train.py:28-52 full width), batch 3, 1024 + 256 steps, random init",
                          cascade_seconds=total, per_stage=stage, output_shape=list(full.shape), dtype="f16",
                          note="north_star names bf16 for this config; the path computes in fp16 with saturating conversions (DESIGN.md section 2)")), flush=True)


if __name__ == "__main__":
    cfg1("--no-cpu" in sys.argv)
    cfg2()
