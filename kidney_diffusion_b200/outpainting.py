"""outpainting.py entry points (reference outpainting.py:173-243): the same wavefront patch sampler on a square
``--num_patches_width`` grid without conditioning images; the stitched canvas starts from zeros.  Models come from
`grid.default_model_provider('outpaint', ...)`: the unconditional cascade of train_uncond.py:79-93 with the checkpoints
args.unet1 / args.unet2 / args.unet3 (outpainting.py:24-45)."""
from __future__ import annotations

import torch

from . import grid
from .grid import PATCH_SIZE


def generate_image_with_unet(unet_number, args, lowres_image, overlap, orientation, num_patches_width):
    """outpainting.py:173-216."""
    patch_pos = [(i, j) for i in range(num_patches_width) for j in range(num_patches_width)]
    return grid.generate_image_with_unet("outpaint", unet_number, args, lowres_image, None, patch_pos, overlap, orientation, num_patches_width)


def generate_image(args, overlap=0.25, orientation=-1, num_patches_width=1):
    """outpainting.py:219-224 (one pipelined plan over the three stages unless args.stage_major)."""
    patch_pos = [(i, j) for i in range(num_patches_width) for j in range(num_patches_width)]
    return grid.generate_image("outpaint", args, cond_image=None, patch_pos=patch_pos, overlap=overlap, orientation=orientation,
                               num_patches_width=num_patches_width)


def generate_high_res_image(args):
    """outpainting.py:227-243: zeros canvas, patches pasted row-major; the image is returned on every rank."""
    n = args.num_patches_width
    patch_pos = [(i, j) for i in range(n) for j in range(n)]
    images = generate_image(args, overlap=args.overlap, orientation=-1, num_patches_width=n)
    dist, rank, world = grid._dist()
    device = grid._device_for(args, rank)
    if device.type == "cuda":
        return grid.stitch_device(None, images, patch_pos, n, args.overlap, device)
    gathered = grid.gather_patches(images)
    full_image = None
    if gathered is not None:
        patch_dist = int(PATCH_SIZE * (1 - args.overlap))
        width = PATCH_SIZE + (n - 1) * patch_dist
        full_image = torch.zeros(1, 3, width, width)
        for index, (i, j) in enumerate(patch_pos):
            full_image[0, :, i * patch_dist:i * patch_dist + PATCH_SIZE, j * patch_dist:j * patch_dist + PATCH_SIZE] = gathered[index][0].cpu()
    if world > 1:
        box = [full_image]
        dist.broadcast_object_list(box, 0)
        full_image = box[0]
    return full_image
