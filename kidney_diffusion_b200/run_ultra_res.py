"""Command-line entry point of the gigapixel sampler with the reference's own flags (sample_ultra_res.py:451-492):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m kidney_diffusion_b200.run_ultra_res \\
        --unet1_mag0 ... --unet3_mag2 ... --inpaint_resample 5 --overlap 0.25 --sample_dir samples --version v_param

One process per GPU (the reference spawns `--num_gpus` worker processes itself; here the launcher does, and `--num_gpus` is
accepted and ignored).  The flow is the reference's main(): magnification 0 image -> 6 400^2 magnification-1 image -> magnification-2
image, each saved as JPEG by rank 0.  Extra flags: --seed (reproducible noise), --stage_major (the reference's stage order instead
of the pipelined plan), --max_mag (stop after this magnification level), --precision fp32 (every UNet on the fp32 CUDA-core path:
validation runs, 20-50x slower).
"""
from __future__ import annotations

import argparse
import os
from uuid import uuid4

import torch


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    for u in (1, 2, 3):
        for m in (0, 1, 2):
            parser.add_argument(f"--unet{u}_mag{m}", type=str)
    parser.add_argument("--num_gpus", type=int)
    parser.add_argument("--inpaint_resample", type=int)
    parser.add_argument("--overlap", type=float)
    parser.add_argument("--sample_dir", default="samples", type=str)
    parser.add_argument("--ignore_unet_1", action="store_true")
    parser.add_argument("--version", type=str)
    parser.add_argument("--seed", type=int, default=None)
    parser.add_argument("--stage_major", action="store_true")
    parser.add_argument("--max_mag", type=int, default=2)
    parser.add_argument("--precision", choices=("fp16", "fp32"), default=None)
    return parser.parse_args(argv)


def main(argv=None, save=None):
    from . import grid

    args = parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    started = False
    if world > 1 and not torch.distributed.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            torch.distributed.init_process_group("gloo")
        started = True
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    if save is None:
        from torchvision.utils import save_image as save
    if rank == 0:
        os.makedirs(args.sample_dir, exist_ok=True)
    sample_id = uuid4()
    postfix = "" if args.version is None or args.version == "" else "-" + args.version
    out = {}
    mag0_images = grid.generate_image(0, args)                                   # sample_ultra_res.py:463
    out[0] = mag0_images[0]
    if rank == 0:
        save(mag0_images[0][0], f"{args.sample_dir}/MAG0-{sample_id}{postfix}.jpg")
    image = mag0_images[0]
    for mag in (1, 2):                                                            # :466-470
        if mag > args.max_mag:
            break
        image = grid.generate_high_res_image(image, mag, args)
        out[mag] = image
        if rank == 0:
            save(image[0], f"{args.sample_dir}/MAG{mag}-{sample_id}{postfix}.jpg")
    if started:
        torch.distributed.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
