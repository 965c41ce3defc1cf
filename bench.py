#!/usr/bin/env python
"""Headline benchmark: "1024^2 ultra-res patch-steps/sec and 16k^2 image time @1/2/4/8 B200" (BASELINE.json metric).

N = 1   value = BASELINE configs[2]: the ultra-res SR UNet 256->1024 of train_ultra_res_v_param.py:51-60 (686 M parameters,
        12.085 TFLOP per sample per step), v-parameterisation, random init, batch 16 of 1024^2 patches.  One *step* = one
        inner iteration of p_sample_loop on the batch: UNet forward (CUDA graph of hand-written kernels) + exact dynamic
        threshold (K7) + fused p_sample update (K6) + on-device Philox noise.
N > 1   value = the same metric through the REAL gigapixel sampler (BASELINE configs[3]): the 1024^2 stage of the 21 x 21 grid
        of overlapping patches (16 384^2 image) -- wavefront dependency DAG, kd_border_pack inpainting canvases, border strips
        between GPUs through the CUDA-IPC peer mailbox -- with `--steps` sampling steps per patch: 441 * steps patch-steps
        divided by the time of the whole stage (max over ranks).  No data-path collective.
every N `grid`: the second half of the metric.  The complete pipeline (get_cond_images -> one pipelined plan over the 64^2 /
        256^2 / 1024^2 stages -> stitch) is RUN with reduced steps and timed; the full-schedule (1024 / 256 / 256 steps) image
        time is the plan's critical-path simulation fed with per-(stage, batch) step times measured in this run, plus the
        measured per-patch overhead (reduced run measured - simulated).  Labelled as an extrapolation.

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
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET3_GFLOP_PER_SAMPLE_1024 = 12085.1  # SURVEY.md appendix B (algorithmic: convs + linears + attention)
METRIC = "ultra_res_1024_patch_steps_per_sec"
UNIT = "patch-steps/s"
FULL_STEPS = {1: 1024, 2: 256, 3: 256}
WORKLOAD_3 = "cfg3 ultra-res SR UNet 256->1024 (train_ultra_res_v_param.py:51-60) v-param patch-step, random init"


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
        self.nvml_rows, self.nvml = [], None
        try:  # NVML (what nvidia-smi itself reads) answers in microseconds: a dense second series for short timed regions, where
            import pynvml  # one nvidia-smi process (0.3-0.8 s each on these boxes) may yield only one or two samples

            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(index))
        except Exception:
            self.nvml = None

    def _nvml_sample(self):
        nv, h = self.nvml
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(reasons_fn(h))
        try:
            watts = nv.nvmlDeviceGetPowerUsage(h) / 1e3
        except Exception:
            watts = None
        self.nvml_rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)), mask, watts))

    def run(self):
        if self.nvml is not None:
            threading.Thread(target=self._nvml_loop, daemon=True).start()
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def _nvml_loop(self):
        while not self.stop_flag.is_set():
            try:
                self._nvml_sample()
            except Exception:
                return
            self.stop_flag.wait(0.02)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = {n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")}
        out = dict(samples=len(self.rows), source="nvidia-smi")
        if self.nvml_rows:  # NVML reason bits: sw_power_cap 0x4, hw_slowdown 0x8, sw_thermal_slowdown 0x20, hw_thermal_slowdown 0x40
            bits = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            nv_sm = [r[0] for r in self.nvml_rows]
            reasons |= {n for r in self.nvml_rows for n, b in bits.items() if r[2] & b}
            out.update(nvml_samples=len(nv_sm), nvml_sm_mhz=statistics.median(nv_sm), nvml_sm_min_mhz=min(nv_sm))
            watts = [r[3] for r in self.nvml_rows if len(r) > 3 and r[3] is not None]
            if watts:
                out["power_w"] = statistics.median(watts)
            if len(sm) < 3:  # too few nvidia-smi samples inside a short region: report the dense series
                sm, mx = nv_sm, [r[1] for r in self.nvml_rows]
                out["source"] = "nvml (fewer than 3 nvidia-smi samples fell inside the timed region)"
        out.update(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons))
        return out


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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    size = args.cpu_size
    # a 1024^2 patch-step costs ~12 TFLOP in fp32 on the host (tens of seconds): the timed sample is bounded to 2 steps + 1 warm-up
    steps, warmup = (min(args.steps, 2), min(args.warmup, 1)) if size >= 1024 else (args.steps, args.warmup)
    times = oracle_patch_step_seconds(size, steps, warmup, threads)
    scale = (1024 / size) ** 2
    ms = 1e3 * sum(times) / len(times) * scale
    value = 1e3 / ms
    sample = (f"{steps} timed + {warmup} warm-up patch-steps (UNet forward + torch.quantile dynamic threshold + posterior update) of the "
              f"config-3/4 SR UNet at {size}x{size}, B=1, fp32 PyTorch eager on {threads} host threads"
              + ("" if size == 1024 else f"; time scaled x{scale:g} (pixel count) to a 1024x1024 patch")
              + (f"; --steps {args.steps} / --warmup {args.warmup} capped to keep the run within minutes" if (steps, warmup) != (args.steps, args.warmup) else ""))
    line = dict(
        impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup, ms_per_step=ms,
        higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        config=dict(workload=WORKLOAD_3, global_batch=1, patch=1024, l2="inputs larger than L2"),
        cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind="port", sample=sample),
        e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
        note="oracle = CPU restatement of imagen-pytorch 1.18.5 (parity unpinned: dependency not installable here)",
    )
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ B200 arm: helpers
class Ctx:
    pass


def make_ctx():
    import torch
    import torch.distributed as dist

    from kidney_diffusion_b200.build import build_library

    c = Ctx()
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback for the product path)")
    torch.cuda.set_device(c.local_rank)
    c.dev = torch.device("cuda", c.local_rank)
    if c.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=c.dev)
    build_library()
    c.peaks = load_peaks()
    c.dist = dist

    def barrier():
        torch.cuda.synchronize()
        if c.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if c.world == 1:
            return ms
        t = torch.tensor([ms], device=c.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    c.barrier, c.max_over_ranks = barrier, max_over_ranks
    return c


def grid_provider(steps_box):
    """MODEL_PROVIDER for the synthetic grid: the three v_param stage models, random init (seeded per stage, identical on every
    rank); steps_box is a mutable {stage: sampling steps} the bench changes between runs."""
    import torch

    from kidney_diffusion_b200.factories import init_imagen_ultra_res, randomize_zero_init_

    def provider(mag, n, device, a):
        torch.manual_seed(10 + n)
        im = init_imagen_ultra_res(mag, n, version="v_param", timesteps=(steps_box[1], steps_box[2], steps_box[3]))
        randomize_zero_init_(im)
        return im.to(device).eval()

    return provider


def set_steps(grid, dev, args, steps):
    for u in (1, 2, 3):
        key = [k for k in grid._MODEL_CACHE if k[0] == 1 and k[1] == u]
        for k in key:
            grid._MODEL_CACHE[k].noise_schedulers[u - 1].num_timesteps = steps[u]


def measure_step_table(c, grid, args, sizes):
    """ms per sampling step of a batch of B patches for every (stage, B) a plan uses: 2 warm-up + 3 timed inner iterations (median)
    of the real StageRun (graph replay + dynamic threshold + update), CUDA events.  Rank 0's numbers are broadcast so that every
    rank plans with the same table."""
    import torch

    from kidney_diffusion_b200.imagen import CounterNoise

    table = {}
    noise = CounterNoise(7, 0)
    for u in sorted(sizes):
        im = grid.load_model(1, u, c.dev, args)
        S = grid.PATCH_SIZES[u]
        table[u] = {}
        saved = im.noise_schedulers[u - 1].num_timesteps
        im.noise_schedulers[u - 1].num_timesteps = 8
        for B in sizes[u]:
            cond = torch.rand(B, 3, 1024, 1024, device=c.dev)
            lowres = torch.randn(B, 3, S, S, device=c.dev) if u > 1 else None
            mask = torch.zeros(B, S, S, device=c.dev)
            mask[:, : S // 4] = 1
            run = im.stage_run(u, (B, 3, S, S), noise=noise, lowres_cond_img=lowres, lowres_noise_level=0.2 if u > 1 else None, cond_images=cond,
                               inpaint_images=torch.rand(B, 3, S, S, device=c.dev), inpaint_masks=mask, inpaint_resample_times=1)
            for k in range(2):
                run.step(k)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            torch.cuda.synchronize()
            ev[0].record()
            for k in range(2, 5):
                run.step(k)
                ev[k - 1].record()
            torch.cuda.synchronize()
            table[u][B] = statistics.median(ev[i].elapsed_time(ev[i + 1]) for i in range(3))  # median: one hiccup must not bend the plan
            del run, cond, lowres
        im.noise_schedulers[u - 1].num_timesteps = saved
    if c.world > 1:
        box = [table]
        c.dist.broadcast_object_list(box, 0)
        table = box[0]
    return table


def grid_section(c, args, steps_box, reduced):
    """Second half of the metric: the 16 384^2 image.  Runs the complete pipeline with `reduced` steps per stage and
    extrapolates to the full schedule through the plan's critical-path simulation with step times measured here."""
    import torch

    from kidney_diffusion_b200 import grid, grid_plan

    n_side = args.grid
    g_args = types.SimpleNamespace(version="v_param", overlap=0.25, inpaint_resample=1, ignore_unet_1=False, num_gpus=c.world, device=None, seed=1234)
    pw = grid.get_patch_width(g_args, 1)
    pd = int(pw * (1 - g_args.overlap))
    W = pw + (n_side - 1) * pd
    zoomed_host = torch.rand(1, 3, W, W, generator=torch.Generator().manual_seed(0)).pin_memory()
    pos = [(i, j) for i in range(n_side) for j in range(n_side)]
    orientation = grid.choose_orientation(pos)
    for u in (1, 2, 3):
        grid.load_model(1, u, c.dev, g_args)

    # 1. plan with the built-in cost model to learn which batch sizes occur, measure those, re-plan with measured times
    full = {u: FULL_STEPS[u] for u in (1, 2, 3)}
    guess = grid_plan.build_plan(pos, orientation, c.world, steps=full)
    sizes = {u: sorted(set(guess.batch_sizes().get(u, [1])) | {1}) for u in (1, 2, 3)}
    t0 = time.time()
    table = measure_step_table(c, grid, g_args, sizes)
    t_table = time.time() - t0
    g_args.cost_table = {u: {**grid_plan.load_cost_table()[u], **table[u]} for u in (1, 2, 3)}  # measured points override the defaults
    plan_full = grid_plan.build_plan(pos, orientation, c.world, steps=full, table=g_args.cost_table)
    g_args.plan_policy = plan_full.policy[0]
    g_args.max_batch = {1: grid.MAX_BATCH[1], 2: grid.MAX_BATCH[2], 3: plan_full.policy[1]}

    # 2. the real pipeline, reduced steps, measured
    set_steps(grid, c.dev, g_args, reduced)
    steps_box.update(reduced)

    def run_once():
        c.barrier()
        t0 = time.time()
        full_img = grid.generate_high_res_image(zoomed_host, 1, g_args)
        host = full_img.cpu() if c.rank == 0 else None
        c.barrier()
        return time.time() - t0, dict(grid.LAST_RUN), full_img, host

    t_cold, _, _, _ = run_once()          # captures CUDA graphs of batch sizes not met yet
    t_warm, info, img, host = run_once()
    finite = bool(torch.isfinite(img).all().item()) and float(img.min()) >= 0.0 and float(img.max()) <= 1.0
    checksum = float(img.double().sum().item())
    shape = list(img.shape)
    del img, host
    t_warm = c.max_over_ranks(t_warm * 1e3) / 1e3
    plan_reduced = grid_plan.build_plan(pos, orientation, c.world, steps=reduced, table=g_args.cost_table, policy=g_args.plan_policy,
                                        max_batch=g_args.max_batch)
    sim_reduced = plan_reduced.makespan
    overhead = max(0.0, info.get("sampling_s", t_warm) - sim_reduced)
    tail = info.get("cond_images_s", 0.0) + info.get("stitch_s", 0.0)
    image_s = plan_full.makespan + overhead + tail
    stage_only = {str(u): grid_plan.build_plan(pos, orientation, c.world, stages=(u,), steps=full, table=g_args.cost_table).makespan for u in (1, 2, 3)}
    set_steps(grid, c.dev, g_args, steps_box)
    return dict(
        workload=f"cfg4: {n_side}x{n_side} grid of overlapping 1024^2 patches ({shape[-1]}^2 image), overlap 0.25, inpaint_resample 1, v_param models, random init",
        image_seconds_extrapolated=image_s,
        how=("critical-path simulation of the full-step plan (1024/256/256 steps; same batches and per-rank order the executor would run) with "
             "the per-(stage,batch) step times measured in this run, + per-patch overhead (measured reduced-step run - its simulation) + measured "
             "cond-image and stitch time"),
        full_plan=dict(makespan_s=plan_full.makespan, busy_fraction=plan_full.busy_fraction(), policy=str(plan_full.policy), batches=len(plan_full.batches),
                       batch_sizes={str(u): v for u, v in plan_full.batch_sizes().items()}),
        per_stage_alone_s=stage_only, stage_major_sum_s=sum(stage_only.values()),
        reduced_run=dict(steps=[reduced[1], reduced[2], reduced[3]], measured_s=t_warm, cold_s=t_cold, sampling_s=info.get("sampling_s"),
                         simulated_sampling_s=sim_reduced, cond_images_s=info.get("cond_images_s"), stitch_s=info.get("stitch_s"),
                         transport=info.get("transport"), bytes_sent_rank0=info.get("bytes_sent_this_rank"), finite_in_0_1=finite, checksum=checksum,
                         image_shape=shape),
        step_ms_measured={str(u): {str(b): round(v, 3) for b, v in table[u].items()} for u in table}, step_table_seconds=t_table,
        patches=len(pos),
    )


def identical_to_1gpu(c, args):
    """Small-grid bit-identity: the N-rank run of the 1024^2 stage (border strips through the peer mailbox) against the same
    grid sampled by rank 0 alone."""
    import torch

    from kidney_diffusion_b200 import grid

    g_args = types.SimpleNamespace(version="v_param", overlap=0.25, inpaint_resample=1, ignore_unet_1=False, num_gpus=c.world, device=None, seed=99,
                                   max_batch=2)
    n_side = 3
    pos = [(i, j) for i in range(n_side) for j in range(n_side)]
    zoomed = torch.rand(1, 3, 166 + 2 * 124, 166 + 2 * 124, generator=torch.Generator().manual_seed(5)).to(c.dev)
    bank, pos2, n = grid.get_cond_images(g_args, zoomed, 1, lazy=True)
    assert n == n_side and pos2 == pos
    lowres = [torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(100 + k)).to(c.dev) for k in range(len(pos))]
    set_steps(grid, c.dev, g_args, {1: 1, 2: 1, 3: 1})
    out_n = grid.generate_image_with_unet(1, 3, g_args, lowres, bank, pos, 0.25, -1, n)
    img_n = grid.stitch_device(zoomed, out_n, pos, n, 0.25, c.dev)
    same = True
    if c.rank == 0:
        grid.DISABLE_DIST = True
        out_1 = grid.generate_image_with_unet(1, 3, g_args, lowres, bank, pos, 0.25, -1, n)
        img_1 = grid.stitch_device(zoomed, out_1, pos, n, 0.25, c.dev)
        grid.DISABLE_DIST = False
        same = bool(torch.equal(img_n, img_1))
    c.barrier()
    flag = torch.tensor([1 if same else 0], device=c.dev)
    if c.world > 1:
        c.dist.broadcast(flag, 0)
    return bool(flag.item())


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch

    from kidney_diffusion_b200 import grid, ops, schedule
    from kidney_diffusion_b200.imagen import CounterNoise

    c = make_ctx()
    world, rank, dev, peaks = c.world, c.rank, c.dev, c.peaks
    B, S, K, W = args.batch, args.size, args.steps, args.warmup
    steps_box = {1: 8, 2: 4, 3: max(K, W)}
    grid.MODEL_PROVIDER = grid_provider(steps_box)
    g_args = types.SimpleNamespace(version="v_param", overlap=0.25, inpaint_resample=1, ignore_unet_1=False, num_gpus=world, device=None, seed=1234)
    imagen = grid.load_model(1, 3, dev, g_args)
    line = dict(metric=METRIC, unit=UNIT, n_gpus=world, steps=K, warmup=W, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16",
                data="synthetic")
    launches0 = ops.launch_count

    if world == 1:
        # ---------------- device-resident throughput (`value`): batch of 16 independent 1024^2 patches, inputs already in HBM
        g = torch.Generator().manual_seed(100 + rank)
        cond_host = torch.rand(B, 3, S, S, generator=g).pin_memory()
        start_host = torch.rand(B, 3, S // 4, S // 4, generator=g).pin_memory()
        out_host = torch.empty(B, 3, S, S).pin_memory()
        noise = CounterNoise(seed=1234, stream_key=rank)
        imagen.noise_schedulers[2].num_timesteps = 256
        cond_dev, start_dev = cond_host.to(dev), start_host.to(dev)
        lowres = imagen.normalize_img(torch.nn.functional.interpolate(start_dev, S, mode="nearest")).contiguous()
        la, ls = schedule.alpha_sigma("linear", 0.2)
        lowres = ops.q_sample(lowres, noise("lowres_aug", tuple(lowres.shape), dev, unet=3), float(la), float(ls))
        run = imagen.stage_run(3, (B, 3, S, S), noise=noise, lowres_cond_img=lowres, lowres_noise_level=0.2, cond_images=cond_dev)
        for k in range(W):
            run.step(k)
        c.barrier()
        sampler = ClockSampler(c.local_rank)
        sampler.start()
        l0 = ops.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(W, W + K):
            run.step(k)
        e1.record()
        c.barrier()
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.summary()
        gpu_launches = ops.launch_count - l0
        ms_per_step = ms_total / K
        value = B * K / (ms_total / 1e3)
        state = run.img
        out_ok = bool(torch.isfinite(state).all().item())
        out_sum = float(state.double().sum().item())
        assert out_ok, "non-finite sampler state after the timed steps"

        # ---------------- roofline of the dominant kernels (tcgen05 conv GEMMs): CUDA events around every launch of one eager step
        ops.conv_profile = []
        imagen.use_cuda_graph = False
        run.step(W + K)
        torch.cuda.synchronize()
        prof, ops.conv_profile = ops.conv_profile, None
        imagen.use_cuda_graph = True
        conv_flops = sum(p[0] for p in prof)
        conv_ms = sum(p[1].elapsed_time(p[2]) for p in prof)
        achieved = conv_flops / (conv_ms / 1e3) / 1e12
        traffic, traffic_note = read_traffic(B)
        roofline = dict(bound="tensor", achieved=achieved, peak=peaks["tf_sustained"], unit="TFLOP/s", frac=achieved / peaks["tf_sustained"],
                        traffic=traffic, traffic_note=traffic_note,
                        kernel="conv_gemm_halo_kernel / conv_gemm_pair_kernel / init_conv_kernel (tcgen05 implicit GEMM, cta_group::2, TMEM)",
                        launches_per_step=len(prof), conv_ms_per_step=conv_ms, conv_share_of_step=conv_ms / ms_per_step,
                        peak_source=peaks["source"] + " sustained bf16",
                        how="sum of algorithmic conv/linear FLOPs of one step / sum of per-launch CUDA-event durations (eager replay of a timed step)")
        unet_tflops = B * UNET3_GFLOP_PER_SAMPLE_1024 * (S / 1024) ** 2 / 1e3 / (ms_per_step / 1e3)
        del run

        # ---------------- live parity at the bench shape: the same patch through the tensor-core path and through the fp32
        # CUDA-core path (csrc/kd_precise.cu, 1e-6 against the fp32 oracle in tests/test_precise_gpu.py); outside every timed region
        unet = imagen.unets[2]
        gp = torch.Generator().manual_seed(7)
        xp = torch.randn(1, 3, S, S, generator=gp).to(dev)
        tp, ltp = torch.tensor([0.9], device=dev), torch.full((1,), float(schedule.log_snr("linear", 0.2)), device=dev)
        kwp = dict(lowres_cond_img=lowres[:1].contiguous(), lowres_noise_times=ltp, cond_images=cond_dev[:1].contiguous())
        fast = unet(xp, tp, **kwp).clone()
        unet.precision = "fp32"
        torch.cuda.synchronize()
        t0 = time.time()
        ref32 = unet(xp, tp, **kwp)
        torch.cuda.synchronize()
        fp32_s = time.time() - t0
        unet.precision = "fp16"
        rn = float(ref32.double().norm().item())
        parity = dict(rel_l2_fp16_path_vs_fp32_path=float((fast.double() - ref32.double()).norm().item()) / rn if rn > 0 else None, tolerance=1e-2,
                      what=f"UNet output of one {S}x{S} patch (B = 1), identical inputs; the fp32 path is itself within 1e-4 of the CPU oracle "
                           "(tests/test_precise_gpu.py, tests/test_unet_parity_gpu.py at 1024^2)", fp32_path_seconds=fp32_s)
        del fast, ref32, xp
        unet._executor = None
        imagen._graphs.clear()
        torch.cuda.empty_cache()

        # ---------------- end to end through the public API: host buffers in, host buffer out, every step
        imagen.noise_schedulers[2].num_timesteps = 1  # one sample() call == one patch-step on the batch (plus the per-call conditioning work)

        def e2e_step():
            out = imagen.sample(batch_size=B, cond_images=cond_host, start_image_or_video=start_host, start_at_unet_number=3,
                                stop_at_unet_number=3, use_tqdm=False, device=dev, noise_key=rank)
            out_host.copy_(out, non_blocking=True)

        for _ in range(W):
            e2e_step()
        c.barrier()
        e0.record()
        for _ in range(K):
            e2e_step()
        e1.record()
        c.barrier()
        e2e_ms = e0.elapsed_time(e1)
        e2e = dict(value=B * K / (e2e_ms / 1e3), unit=UNIT, h2d_bytes_per_step=cond_host.numel() * 4 + start_host.numel() * 4,
                   d2h_bytes_per_step=out_host.numel() * 4, ms_per_step=e2e_ms / K,
                   call="Imagen.sample(cond_images=<pinned host>, start_image_or_video=<pinned host>, start/stop_at_unet_number=3) with a "
                        "1-step schedule + copy of the result to pinned host memory")
        del cond_dev, start_dev, lowres, state
        imagen._graphs.clear()
        torch.cuda.empty_cache()
        line.update(value=value, ms_per_step=ms_per_step,
                    config=dict(workload=WORKLOAD_3, global_batch=B, per_gpu_batch=B, patch=S, parallelism="1 GPU, batch of 16 independent patches",
                                l2="inputs larger than L2 (activations >= 0.27 GB per tensor per patch)", state_dtype="f32",
                                unet_gflop_per_patch_step=UNET3_GFLOP_PER_SAMPLE_1024 * (S / 1024) ** 2),
                    roofline=roofline, e2e=e2e, gpu_launches=gpu_launches, clocks=clocks, unet_algorithmic_tflops=unet_tflops,
                    tensor_frac_whole_step=unet_tflops / peaks["tf_sustained"], output=dict(finite=out_ok, checksum=out_sum), parity=parity)
    else:
        # ---------------- N > 1: the 1024^2 stage of the real 21 x 21 wavefront grid
        n_side = args.grid
        pw = grid.get_patch_width(g_args, 1)
        pd = int(pw * (1 - g_args.overlap))
        Wz = pw + (n_side - 1) * pd
        zoomed_host = torch.rand(1, 3, Wz, Wz, generator=torch.Generator().manual_seed(0)).pin_memory()
        zoomed = zoomed_host.to(dev)
        bank, pos, n = grid.get_cond_images(g_args, zoomed, 1, lazy=True)
        assert n == n_side
        orientation = grid.choose_orientation(pos)
        lowres_host = torch.rand(len(pos), 3, 256, 256, generator=torch.Generator().manual_seed(3)).pin_memory()
        lowres_dev = lowres_host.to(dev)
        lowres = [lowres_dev[k:k + 1] for k in range(len(pos))]

        def stage3(steps, lr, bnk):
            set_steps(grid, dev, g_args, {1: steps_box[1], 2: steps_box[2], 3: steps})
            return grid.generate_image_with_unet(1, 3, g_args, lr, bnk, pos, g_args.overlap, orientation, n)

        stage3(W, lowres, bank)  # warm-up: W sampling steps of the whole grid stage (captures every batch size's CUDA graph)
        c.barrier()
        sampler = ClockSampler(c.local_rank)
        sampler.start()
        l0 = ops.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = stage3(K, lowres, bank)
        e1.record()
        c.barrier()
        ms_total = c.max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.summary()
        gpu_launches = ops.launch_count - l0
        info = dict(grid.LAST_RUN)
        value = len(pos) * K / (ms_total / 1e3)
        mine = [p for p in out if p is not None]
        ok = torch.tensor([1 if all(bool(torch.isfinite(p).all().item()) for p in mine) else 0], device=dev)
        csum = torch.tensor([sum(float(p.double().sum().item()) for p in mine)], device=dev, dtype=torch.float64)
        c.dist.all_reduce(ok, op=c.dist.ReduceOp.MIN)
        c.dist.all_reduce(csum)
        assert int(ok.item()) == 1, "non-finite patch in the grid stage"
        plan = out.plan
        sent = torch.tensor([float(info.get("bytes_sent_this_rank", 0))], device=dev, dtype=torch.float64)
        c.dist.all_reduce(sent)

        # e2e: same stage through the public entry points with HOST inputs and a HOST result
        canvas_host = torch.empty(1, 3, 1024 + (n - 1) * 768, 1024 + (n - 1) * 768).pin_memory() if rank == 0 else None

        def e2e_run(steps):
            bnk = grid.get_cond_images(g_args, zoomed_host.to(dev, non_blocking=True), 1, lazy=True)[0]
            lr = [lowres_host[k:k + 1] for k in range(len(pos))]  # each patch's low-res input is uploaded by the rank that samples it
            o = stage3(steps, lr, bnk)
            full = grid.stitch_device(zoomed, o, pos, n, g_args.overlap, dev)
            if rank == 0:
                canvas_host.copy_(full, non_blocking=True)
            return full

        c.barrier()
        e0.record()
        full = e2e_run(K)
        e1.record()
        c.barrier()
        e2e_ms = c.max_over_ranks(e0.elapsed_time(e1))
        e2e = dict(value=len(pos) * K / (e2e_ms / 1e3), unit=UNIT, h2d_bytes_per_step=(zoomed_host.numel() * 4 * world + lowres_host.numel() * 4) // K,
                   d2h_bytes_per_step=full.numel() * 4 // K, ms_per_step=e2e_ms / K,
                   call="get_cond_images(<pinned host image>) + generate_image_with_unet(1, 3, ..., <pinned host low-res patches>) + stitch_device + "
                        "copy of the stitched 16 384^2 image to pinned host memory; bytes are per sampling step of the whole grid stage")
        del full
        line.update(value=value, ms_per_step=ms_total / K,
                    config=dict(workload=f"cfg4 stage 3: {WORKLOAD_3}, sampled as the 1024^2 stage of the {n_side}x{n_side} overlapping-patch grid "
                                         f"(16 384^2 image) with {K} sampling steps per patch",
                                global_batch=len(pos), patch=1024, parallelism="wavefront DAG over patches, border exchange through CUDA-IPC peer mailbox (NVLink), no collective",
                                l2="inputs larger than L2 (activations >= 0.27 GB per tensor per patch)", state_dtype="f32",
                                plan=dict(batches=len(plan.batches), batch_sizes=plan.batch_sizes().get(3), policy=str(plan.policy),
                                          simulated_busy_fraction=plan.busy_fraction()),
                                transport=info.get("transport"), border_bytes_exchanged=float(sent.item()), trace_rank0=info.get("trace"),
                                unet_gflop_per_patch_step=UNET3_GFLOP_PER_SAMPLE_1024),
                    roofline=dict(bound="tensor", achieved=value * UNET3_GFLOP_PER_SAMPLE_1024 / 1e3 / world, peak=peaks["tf_sustained"], unit="TFLOP/s",
                                  frac=value * UNET3_GFLOP_PER_SAMPLE_1024 / 1e3 / world / peaks["tf_sustained"], traffic=None,
                                  kernel="whole patch-step per GPU (dominant kernels: tcgen05 conv GEMMs; per-kernel roofline is reported by the N=1 run)",
                                  how="algorithmic UNet FLOPs of the patch-steps all ranks processed / (ranks x stage time): includes the wavefront's idle time",
                                  peak_source=peaks["source"] + " sustained bf16"),
                    e2e=e2e, gpu_launches=gpu_launches, clocks=clocks, output=dict(finite=True, checksum=float(csum.item())),
                    cpu_baseline=None)
        del lowres, lowres_dev, out
        torch.cuda.empty_cache()
        line["identical_to_1gpu"] = identical_to_1gpu(c, args)

    # ---------------- second half of the metric: the 16k^2 image through the complete pipeline
    if not args.no_grid:
        reduced = {1: args.grid_steps[0], 2: args.grid_steps[1], 3: args.grid_steps[2]}
        line["grid"] = grid_section(c, args, steps_box, reduced)
        if world > 1:
            line["grid"]["identical_to_1gpu"] = line.get("identical_to_1gpu")
    line["gpu_launches_total"] = ops.launch_count - launches0

    # ---------------- CPU baseline beside it (rank 0, N = 1 only)
    if world == 1:
        cpu_baseline = None
        if not args.no_cpu_baseline:
            grid._MODEL_CACHE.clear()
            imagen = None
            torch.cuda.empty_cache()
            threads = os.cpu_count() or 1
            size = args.cpu_size
            times = oracle_patch_step_seconds(size, 1, 0 if size >= 1024 else 1, threads)
            scale = (1024 / size) ** 2
            sec = sum(times) / len(times) * scale
            cpu_baseline = dict(value=1.0 / sec, unit=UNIT, cores=threads, kind="port",
                                sample=f"1 patch-step (UNet forward + torch.quantile dynamic threshold + posterior update) of the same config-3/4 UNet at "
                                       f"{size}x{size}, B=1, fp32 PyTorch CPU oracle, {threads} threads"
                                       + ("" if size == 1024 else f"; time scaled x{scale:g} (pixel count) to a 1024x1024 patch"))
        line["cpu_baseline"] = cpu_baseline

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        c.dist.destroy_process_group()
    return 0


def read_traffic(B):
    """DRAM bytes of the dominant launch shape from the committed `ncu --set full` capture of THIS build (profiles/r02_traffic.json carries
    the build stamp of the library it was captured from; a stale capture is reported as null), scaled from the captured batch."""
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    spath = os.path.join(ROOT, "kidney_diffusion_b200", ".build_stamp")
    if not os.path.exists(tpath):
        return None, "no ncu traffic capture committed for this build"
    with open(tpath) as f:
        tj = json.load(f)
    stamp = open(spath).read().strip() if os.path.exists(spath) else None
    conv_hash = tj.get("conv_source_sha256")
    if conv_hash:
        import hashlib

        cur = hashlib.sha256(open(os.path.join(ROOT, "kidney_diffusion_b200", "csrc", "kd_conv_gemm.cu"), "rb").read()).hexdigest()
        if cur != conv_hash:
            return None, f"profiles/r02_traffic.json was captured from another version of kd_conv_gemm.cu (stale): not reported"
    traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * B / tj["batch"]
    note = (f"{tj['kernel']} on {tj['shape']}: ncu dram read+write {tj['dram_bytes_read'] + tj['dram_bytes_write']:.3e} B at B={tj['batch']} "
            f"(algorithmic {tj['algorithmic_bytes']:.3e} B), scaled to B={B}; source {tj['source']}; library stamp {stamp}")
    return traffic, note


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="1024^2 patches per step at N = 1 (BASELINE.json configs[2]: 16)")
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--grid", type=int, default=21, help="patches per side of the gigapixel grid (21 -> 16 384^2 at overlap 0.25)")
    ap.add_argument("--grid-steps", default="8,2,2", help="reduced sampling steps (64^2, 256^2, 1024^2 stages) of the measured full-pipeline run")
    ap.add_argument("--no-grid", action="store_true", help="skip the 16k^2 image section")
    ap.add_argument("--cpu-size", type=int, default=1024, help="patch size of the CPU sample (1024 = the real patch, no scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.grid_steps = tuple(int(v) for v in args.grid_steps.split(","))
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
