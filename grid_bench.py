#!/usr/bin/env python
"""Second headline number of BASELINE.json: time of a 16 384^2 ultra-res image (21 x 21 grid of overlapping 1024^2 patches,
config 4) on N B200s of one node.

The full schedule (1024 / 256 / 256 steps per stage, 441 patches) is 1.5 EFLOP -- minutes even at roofline -- so this
driver runs the REAL pipeline (get_cond_images -> wavefront schedule -> per-stage sampling with RePaint inpainting ->
NCCL border exchange -> gather -> stitch) with a reduced number of steps per stage and reports (a) the measured
reduced-step time per stage and (b) the full-schedule image time obtained by scaling each stage's measured sampling time by
full_steps / reduced_steps (labelled as an extrapolation; setup, exchange and stitch times are taken as measured).

  python grid_bench.py --steps 4,2,2                         # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 grid_bench.py --steps 4,2,2
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL_STEPS = (1024, 256, 256)


def main():
    import torch
    import torch.distributed as dist

    from kidney_diffusion_b200 import grid, ops
    from kidney_diffusion_b200.build import build_library
    from kidney_diffusion_b200.factories import init_imagen_ultra_res, randomize_zero_init_

    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", default="4,2,2", help="reduced sampling steps for the 64 / 256 / 1024 stages")
    ap.add_argument("--grid", type=int, default=21, help="patches per side (21 -> 16 384^2 at overlap 0.25)")
    ap.add_argument("--resample", type=int, default=1, help="--inpaint_resample of the reference (RePaint inner iterations)")
    ap.add_argument("--overlap", type=float, default=0.25)
    args_cli = ap.parse_args()
    steps = tuple(int(s) for s in args_cli.steps.split(","))

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    build_library()

    def provider(mag, n, device, a):
        torch.manual_seed(10 + n)
        im = init_imagen_ultra_res(mag, n, version="v_param", timesteps=steps)
        randomize_zero_init_(im)
        return im.to(device).eval()

    grid.MODEL_PROVIDER = provider
    args = types.SimpleNamespace(version="v_param", overlap=args_cli.overlap, inpaint_resample=args_cli.resample, ignore_unet_1=False,
                                 num_gpus=world, device=None)
    # zoomed image whose mag-1 grid has exactly `--grid` patches per side: n = 1 + ceil((W - 166) / 124)
    pw = grid.get_patch_width(args, 1)
    pd = int(pw * (1 - args.overlap))
    W = pw + (args_cli.grid - 1) * pd
    zoomed = torch.rand(1, 3, W, W, generator=torch.Generator().manual_seed(0))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.time()
    cond, pos, n = grid.get_cond_images(args, zoomed, 1)
    assert n == args_cli.grid, (n, args_cli.grid)
    orientation = grid.choose_orientation(pos)
    t_cond = time.time() - t0
    for u in (1, 2, 3):
        grid.load_model(1, u, dev, args)  # build + upload all three stage models before timing (the reference reloads per stage)
    sync()

    def set_steps(st):
        for u in (1, 2, 3):
            grid.load_model(1, u, dev, args).noise_schedulers[u - 1].num_timesteps = st[u - 1]

    def run_pipeline(st):
        set_steps(st)
        stage_t, prev = {}, None
        t_all = time.time()
        for u in (1, 2, 3):
            sync()
            t0 = time.time()
            prev = grid.generate_image_with_unet(1, u, args, prev, cond, pos, args.overlap, orientation, n)
            sync()
            stage_t[u] = time.time() - t0
        t0 = time.time()
        patches = grid.gather_patches(prev)
        sync()
        t_gather = time.time() - t0
        t_stitch, shape = 0.0, None
        if rank == 0:
            t0 = time.time()
            full = grid.stitch(zoomed, patches, pos, n, args.overlap)
            t_stitch = time.time() - t0
            shape = list(full.shape)
            assert float(full.min()) >= 0.0 and float(full.max()) <= 1.0
        return dict(total_s=time.time() - t_all, stage_s=stage_t, gather_s=t_gather, stitch_s=t_stitch), shape

    rounds = {u: len(grid.build_schedule(pos, orientation, world, grid.MAX_BATCH[u]).rounds) for u in (1, 2, 3)}
    launches0 = ops.launch_count
    steps2 = tuple(2 * s for s in steps)
    cold, shape = run_pipeline(steps)      # includes one-time CUDA-graph captures and allocator warm-up
    warm1, _ = run_pipeline(steps)
    warm2, _ = run_pipeline(steps2)
    if rank == 0:
        # per-stage linear model t = fixed + slope * steps, slope from the two warm runs, fixed part from the cold run
        slope = {u: max(0.0, (warm2["stage_s"][u] - warm1["stage_s"][u]) / (steps2[u - 1] - steps[u - 1])) for u in (1, 2, 3)}
        fixed = {u: max(0.0, cold["stage_s"][u] - slope[u] * steps[u - 1]) for u in (1, 2, 3)}
        full_stage = {u: fixed[u] + slope[u] * FULL_STEPS[u - 1] * 1.0 for u in (1, 2, 3)}
        extrap = t_cond + sum(full_stage.values()) + cold["gather_s"] + cold["stitch_s"]
        per_patch_step_ms = {u: 1e3 * slope[u] / (len(pos) * args_cli.resample) for u in (1, 2, 3)}
        line = dict(
            metric="ultra_res_16k_image_seconds", unit="s", higher_is_better=False, n_gpus=world, data="synthetic", dtype="f16",
            config=dict(workload=f"cfg4 {n}x{n} grid of overlapping 1024^2 patches ({shape[-1] if shape else '?'}^2 image), overlap {args.overlap}, "
                                 f"inpaint_resample {args_cli.resample}, v_param models, random init", reduced_steps=[list(steps), list(steps2)],
                        full_steps=list(FULL_STEPS), max_batch=grid.MAX_BATCH, rounds_per_stage=rounds, patches=len(pos)),
            measured_reduced=dict(cold=cold, warm=warm1, warm_double_steps=warm2, cond_images_s=t_cond),
            model=dict(seconds_per_sampling_step_of_the_whole_grid=slope, fixed_seconds=fixed, full_stage_seconds=full_stage,
                       amortised_ms_per_patch_step=per_patch_step_ms),
            value=extrap, value_kind="extrapolated: per stage fixed + slope * full steps (slope from two warm reduced-step runs of the real "
                                     "pipeline, fixed part from the cold run) + measured cond-image build, gather and stitch",
            gpu_launches=ops.launch_count - launches0, image_shape=shape,
        )
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
