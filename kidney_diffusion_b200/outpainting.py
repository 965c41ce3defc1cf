"""outpainting.py entry points (reference outpainting.py:173-243): the same wavefront patch sampler on a square
``--num_patches_width`` grid without conditioning images; the stitched canvas starts from zeros."""
from __future__ import annotations

import torch

from . import grid
from .grid import PATCH_SIZE


def default_model_provider(mag_level, unet_number, device, args):
    """outpainting.py load_model: the unconditional cascade (train_uncond.py:79-93) + checkpoint args.unet{n}."""
    from .factories import init_imagen_uncond
    from .trainer import restore_parts

    imagen = init_imagen_uncond(unet_number, device=device)
    loaded = torch.load(vars(args)[f"unet{unet_number}"], map_location="cpu")
    try:
        imagen.load_state_dict(loaded["model"], strict=True)
    except RuntimeError:
        imagen.load_state_dict(restore_parts(imagen.state_dict(), loaded["model"]))
    return imagen


def generate_image_with_unet(unet_number, args, lowres_image, overlap, orientation, num_patches_width):
    """outpainting.py:173-216."""
    patch_pos = [(i, j) for i in range(num_patches_width) for j in range(num_patches_width)]
    return grid.generate_image_with_unet("outpaint", unet_number, args, lowres_image, None, patch_pos, overlap, orientation, num_patches_width)


def generate_image(args, overlap=0.25, orientation=-1, num_patches_width=1):
    """outpainting.py:219-224."""
    low = generate_image_with_unet(1, args, None, overlap, orientation, num_patches_width)
    med = generate_image_with_unet(2, args, low, overlap, orientation, num_patches_width)
    return generate_image_with_unet(3, args, med, overlap, orientation, num_patches_width)


def generate_high_res_image(args):
    """outpainting.py:227-243: zeros canvas, patches pasted row-major (rank 0 gets the image, other ranks None)."""
    n = args.num_patches_width
    images = grid.gather_patches(generate_image(args, overlap=args.overlap, orientation=-1, num_patches_width=n))
    if images is None:
        return None
    patch_dist = int(PATCH_SIZE * (1 - args.overlap))
    width = PATCH_SIZE + (n - 1) * patch_dist
    full_image = torch.zeros(1, 3, width, width)
    for index, (i, j) in enumerate((i, j) for i in range(n) for j in range(n)):
        full_image[0, :, i * patch_dist:i * patch_dist + PATCH_SIZE, j * patch_dist:j * patch_dist + PATCH_SIZE] = images[index][0].cpu()
    return full_image
