#!/usr/bin/env python
"""Headline benchmark: 1024^2 ultra-res patch-steps/sec (BASELINE.json metric) on N B200s of one node.

Workload (BASELINE.json configs[2]): the ultra-res SR UNet 256->1024 of train_ultra_res_v_param.py:51-60 (686 M parameters,
12.085 TFLOP per sample per step), v-parameterisation, random init, batch 16 of 1024^2 patches per GPU.  One *step* =
one inner iteration of p_sample_loop on the batch: UNet forward (CUDA graph of hand-written kernels) + exact dynamic
threshold (K7) + fused p_sample update (K6) + on-device Philox noise.  N > 1: every rank runs its own batch of patches
(patches are independent units; no data-path collective) -> weak scaling.

  python bench.py --gpus N --steps K --warmup W            # B200 arm, one JSON line on rank 0
  python bench.py --impl reference --steps K --warmup W    # the reference path (fp32 PyTorch CPU oracle) on host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET3_GFLOP_PER_SAMPLE_1024 = 12085.1  # SURVEY.md appendix B (algorithmic: convs + linears + attention)
METRIC = "ultra_res_1024_patch_steps_per_sec"
UNIT = "patch-steps/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows))


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def oracle_patch_step_seconds(size, steps, warmup, threads):
    """Times the reference path's patch-step (fp32 PyTorch eager on CPU: UNet + torch.quantile + posterior update) with
    the oracle restatement of imagen-pytorch 1.18.5 (the package itself cannot be installed here: no network / wheel)."""
    import torch

    from oracle import imagen_oracle as O

    torch.set_num_threads(threads)
    torch.manual_seed(0)
    kw = dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 6, 8), memory_efficient=True, layer_attns=False,
              layer_cross_attns=(False, False, False, True), init_conv_to_final_conv_residual=True, cond_images_channels=3)
    null1, null2 = O.NullUnet(), O.NullUnet()
    null2.lowres_cond = True
    imagen = O.Imagen(unets=(null1, null2, O.Unet(**kw)), image_sizes=(size // 16, size // 4, size), timesteps=(1024, 256, 256),
                      pred_objectives=("noise", "v", "v"), condition_on_text=False).eval()
    unet, sched = imagen.unets[2], imagen.noise_schedulers[2]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 3, size, size, generator=g)
    lowres = torch.randn(1, 3, size, size, generator=g)
    cond = torch.rand(1, 3, size, size, generator=g)
    lt = torch.full((1,), 0.2)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t = torch.full((1,), 1.0 - i / 256)
            t0 = time.perf_counter()
            x, _ = imagen.p_sample(unet, x, t, t_next=t - 1 / 256, noise=torch.randn(x.shape, generator=g), noise_scheduler=sched,
                                   cond_images=cond, lowres_cond_img=lowres, lowres_noise_times=lt, pred_objective="v")
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    size = args.cpu_size
    times = oracle_patch_step_seconds(size, args.steps, args.warmup, threads)
    scale = (1024 / size) ** 2  # the UNet is fully convolutional: FLOPs per patch-step scale with the pixel count
    ms = 1e3 * sum(times) / len(times) * scale
    value = 1e3 / ms
    sample = (f"{args.steps} timed + {args.warmup} warm-up patch-steps of the same UNet/update at {size}x{size}, B=1, fp32 CPU; "
              f"time scaled x{scale:g} (pixel count) to the 1024x1024 patch")
    line = dict(
        impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload="cfg3 ultra-res SR UNet 256->1024 (train_ultra_res_v_param.py:51-60) v-param patch-step, random init",
                    global_batch=1, patch=1024, l2="inputs larger than L2"),
        cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample),
        e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
        note="oracle = CPU restatement of imagen-pytorch 1.18.5 (parity unpinned: dependency not installable here)",
    )
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from kidney_diffusion_b200 import ops
    from kidney_diffusion_b200.build import build_library
    from kidney_diffusion_b200.factories import init_imagen_ultra_res, randomize_zero_init_
    from kidney_diffusion_b200.imagen import CounterNoise

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    build_library()
    peaks = load_peaks()
    B, S, K, W = args.batch, args.size, args.steps, args.warmup

    torch.manual_seed(0)
    imagen = init_imagen_ultra_res(1, 3, version="v_param")
    randomize_zero_init_(imagen)
    imagen = imagen.to(dev).eval()
    g = torch.Generator().manual_seed(100 + rank)
    cond_host = torch.rand(B, 3, S, S, generator=g).pin_memory()
    start_host = torch.rand(B, 3, S // 4, S // 4, generator=g).pin_memory()
    out_host = torch.empty(B, 3, S, S).pin_memory()
    noise = CounterNoise(seed=1234, stream_key=rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput (`value`): inputs already in HBM
    cond_dev, start_dev = cond_host.to(dev), start_host.to(dev)
    lowres = imagen.normalize_img(torch.nn.functional.interpolate(start_dev, S, mode="nearest")).contiguous()
    from kidney_diffusion_b200 import schedule

    la, ls = schedule.alpha_sigma("linear", 0.2)
    lowres = ops.q_sample(lowres, noise("lowres_aug", tuple(lowres.shape), dev, unet=3), float(la), float(ls))
    run = imagen.stage_run(3, (B, 3, S, S), noise=noise, lowres_cond_img=lowres, lowres_noise_level=0.2, cond_images=cond_dev)
    assert W + K <= run.num_steps
    for k in range(W):
        run.step(k)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(W, W + K):
        run.step(k)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.summary()
    gpu_launches = ops.launch_count - launches0
    ms_per_step = ms_total / K
    value = world * B * K / (ms_total / 1e3)

    # ---------------- roofline of the dominant kernel (conv_gemm): CUDA events around every launch of one eager step
    ops.conv_profile = []
    imagen.use_cuda_graph = False
    run.step(W + K - 1)
    torch.cuda.synchronize()
    prof, ops.conv_profile = ops.conv_profile, None
    imagen.use_cuda_graph = True
    conv_flops = sum(p[0] for p in prof)
    conv_ms = sum(p[1].elapsed_time(p[2]) for p in prof)
    achieved = conv_flops / (conv_ms / 1e3) / 1e12
    # DRAM bytes of the dominant launch shape from the committed `ncu --set full` capture (profiles/r01_traffic.json), scaled
    # from the captured batch to this run's: the kernel moves its algorithmic bytes once (no re-reads)
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * B / tj["batch"]
        traffic_note = (f"{tj['kernel']} on {tj['shape']}: ncu dram read+write {tj['dram_bytes_read'] + tj['dram_bytes_write']:.3e} B at B={tj['batch']} "
                        f"(algorithmic {tj['algorithmic_bytes']:.3e} B), scaled to B={B}; source {tj['source']}")
    roofline = dict(bound="tensor", achieved=achieved, peak=peaks["tf_sustained"], unit="TFLOP/s", frac=achieved / peaks["tf_sustained"],
                    traffic=traffic, traffic_note=traffic_note,
                    kernel="conv_gemm_halo_kernel / conv_gemm_pair_kernel / init_conv_kernel (tcgen05 implicit GEMM, cta_group::2, TMEM)",
                    launches_per_step=len(prof),
                    conv_ms_per_step=conv_ms, conv_share_of_step=conv_ms / ms_per_step, peak_source=peaks["source"] + " sustained bf16",
                    how="sum of algorithmic conv/linear FLOPs of one step / sum of per-launch CUDA-event durations (eager replay of a timed step)")
    unet_tflops = B * UNET3_GFLOP_PER_SAMPLE_1024 * (S / 1024) ** 2 / 1e3 / (ms_per_step / 1e3)

    # ---------------- end to end through the public API: host buffers in, host buffer out, every step
    spec = imagen.noise_schedulers[2]
    saved_T = spec.num_timesteps
    spec.num_timesteps = 1  # one sample() call == one patch-step on the batch (plus the per-call conditioning work)
    def e2e_step():
        out = imagen.sample(batch_size=B, cond_images=cond_host, start_image_or_video=start_host, start_at_unet_number=3,
                            stop_at_unet_number=3, use_tqdm=False, device=dev, noise_key=rank)
        out_host.copy_(out, non_blocking=True)
    for _ in range(W):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(K):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    spec.num_timesteps = saved_T
    e2e = dict(value=world * B * K / (e2e_ms / 1e3), unit=UNIT, h2d_bytes_per_step=cond_host.numel() * 4 + start_host.numel() * 4,
               d2h_bytes_per_step=out_host.numel() * 4, ms_per_step=e2e_ms / K,
               call="Imagen.sample(cond_images=<pinned host>, start_image_or_video=<pinned host>, start/stop_at_unet_number=3) with a "
                    "1-step schedule + copy of the result to pinned host memory")

    # ---------------- CPU baseline beside it (rank 0, N = 1 only)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        del run
        torch.cuda.empty_cache()
        threads = os.cpu_count() or 1
        times = oracle_patch_step_seconds(args.cpu_size, 1, 1 if args.cpu_size <= 256 else 0, threads)
        scale = (1024 / args.cpu_size) ** 2
        sec = sum(times) / len(times) * scale
        cpu_baseline = dict(value=1.0 / sec, unit=UNIT, cores=threads, kind="port",
                            sample=f"1 patch-step of the same UNet + quantile + update at {args.cpu_size}x{args.cpu_size}, B=1, fp32 PyTorch CPU "
                                   f"oracle; time scaled x{scale:g} (pixel count) to a 1024x1024 patch")

    if rank == 0:
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_per_step, higher_is_better=True,
            scaling="weak", vs_baseline=None, dtype="f16", data="synthetic",
            config=dict(workload="cfg3 ultra-res SR UNet 256->1024 (train_ultra_res_v_param.py:51-60) v-param patch-step, random init",
                        global_batch=world * B, per_gpu_batch=B, patch=S, parallelism=f"patch-parallel x{world} (no collective)",
                        l2="inputs larger than L2 (activations >= 0.27 GB per tensor per patch)", state_dtype="f32",
                        unet_gflop_per_patch_step=UNET3_GFLOP_PER_SAMPLE_1024 * (S / 1024) ** 2),
            roofline=roofline, cpu_baseline=cpu_baseline, e2e=e2e, gpu_launches=gpu_launches, clocks=clocks,
            unet_algorithmic_tflops=unet_tflops, tensor_frac_whole_step=unet_tflops / peaks["tf_sustained"],
        )
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="1024^2 patches per GPU per step (BASELINE.json configs[2]: 16)")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--cpu-size", type=int, default=512, help="patch size of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
