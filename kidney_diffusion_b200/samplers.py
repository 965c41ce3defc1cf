"""Single-GPU cascade wrappers of the reference (scope row a16 of SURVEY.md section 8): `generate_images` of sample_uncond.py:21-74
and of sample_cond.py:24-56, with the same arguments and behaviour -- build the stage's Imagen, wrap it in ImagenTrainer, load
the stage checkpoint (EMA weights folded in, as trainer.sample uses them), sample in chunks of BATCH_SIZES, hand images on as CPU
tensors between stages (PIL images from the last one).  `init_imagen` defaults to the restated factories and may be replaced by
the reference's own function.
"""
from __future__ import annotations

import gc
from uuid import uuid4

import torch

from .trainer import ImagenTrainer

BATCH_SIZES = [64, 64, 6]  # sample_uncond.py:19


def _checkpoint(args, unet_number):
    return {1: args.unet1_checkpoint, 2: args.unet2_checkpoint}.get(unet_number, args.unet3_checkpoint)


def generate_images_uncond(unet_number, args, lowres_images=None, init_imagen=None, batch_sizes=None):
    """sample_uncond.py:21-74."""
    if init_imagen is None:
        from .factories import init_imagen_uncond

        init_imagen = lambda n: init_imagen_uncond(n).cuda()  # noqa: E731
    imagen = init_imagen(unet_number)
    trainer = ImagenTrainer(imagen=imagen)
    trainer.load(_checkpoint(args, unet_number))
    batch_size = (batch_sizes or BATCH_SIZES)[unet_number - 1]
    all_images = []
    for start_idx in range(0, args.num_images, batch_size):
        end_idx = min(start_idx + batch_size, args.num_images)
        batch_lowres_images = None if lowres_images is None else lowres_images[start_idx:end_idx]
        images = trainer.sample(batch_size=end_idx - start_idx, return_pil_images=(unet_number == 3), start_image_or_video=batch_lowres_images,
                                start_at_unet_number=unet_number, stop_at_unet_number=unet_number)
        if unet_number != 3:
            all_images.append(images.cpu())
        else:
            for image in images:
                image.save(f"{args.folder_name}/inference-{uuid4()}.png")
    del trainer, imagen
    gc.collect()
    torch.cuda.empty_cache()
    if unet_number != 3:
        return torch.cat(all_images, dim=0)


def generate_images_cond(unet_number, args, deep_labelmap, num_variants, lowres_images=None, init_imagen=None):
    """sample_cond.py:24-56: `num_variants` samples for one 4-channel label map and the clinical vector [0.0, 0.5, 0.2]."""
    if init_imagen is None:
        from .factories import init_imagen_cond

        init_imagen = lambda n: init_imagen_cond(n).cuda()  # noqa: E731
    imagen = init_imagen(unet_number)
    trainer = ImagenTrainer(imagen=imagen)
    trainer.load(_checkpoint(args, unet_number))
    conds = torch.tensor([0.0, 0.5, 0.2]).reshape(1, 1, 3).repeat_interleave(num_variants, dim=0).float().cuda()
    deep_labelmap = torch.as_tensor(deep_labelmap).unsqueeze(0).repeat_interleave(num_variants, dim=0).float().cuda()
    images = trainer.sample(batch_size=num_variants, return_pil_images=(unet_number == 3), text_embeds=conds, start_image_or_video=lowres_images,
                            cond_images=deep_labelmap, start_at_unet_number=unet_number, stop_at_unet_number=unet_number)
    del trainer, imagen, conds
    gc.collect()
    torch.cuda.empty_cache()
    return images
