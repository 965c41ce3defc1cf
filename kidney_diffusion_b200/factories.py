"""Model factories with the reference's constructor arguments (scope rows a1-a3 of SURVEY.md section 8).

The reference defines ``unet_generator`` / ``FixedNullUnet`` / ``init_imagen`` once per training script; the scripts cannot
be imported here (their dataset modules need slideio / h5py), so the argument sets are restated with citations:

* ultra-res v_param : train_ultra_res_v_param.py:27-92   (``--version v_param``)
* ultra-res default : train_ultra_res.py:27-92           (unet1 dim_mults (1,2,4,8), all-"noise" objectives)
* ultra-res v2      : train_ultra_res_v2.py:27-92        (cond_images_channels = 6)
* ultra-res airs    : train_ultra_res_airs.py:23-88      (("v","v","v") objectives)
* unconditional     : train_uncond.py:28-93
* mask-conditioned  : train.py:28-112                    (text_embed_dim 3, cond_images_channels 4, cond_dim 512)
"""
from __future__ import annotations

import torch
from torch import nn

from .imagen import Imagen
from .unet import NullUnet, Unet


class FixedNullUnet(NullUnet):
    """train_ultra_res_v_param.py:65-75 (identical in every train*.py)."""

    def __init__(self, lowres_cond=False, *args, **kwargs):
        super().__init__()
        self.lowres_cond = lowres_cond
        self.dummy_parameter = nn.Parameter(torch.tensor([0.0]))

    def cast_model_parameters(self, *args, **kwargs):
        return self

    def forward(self, x, *args, **kwargs):
        return x


def ultra_res_unet(magnification_level, unet_number, version="v_param", **overrides):
    cond_ch = {"v2": 6}.get(version, 3) if magnification_level > 0 else 0
    if unet_number == 1:
        mults = (1, 2, 3, 4) if version in ("v_param", "airs") else (1, 2, 4, 8)
        kw = dict(dim=256, dim_mults=mults, num_resnet_blocks=3, layer_attns=(False, True, True, True),
                  layer_cross_attns=(False, True, True, True), cond_images_channels=cond_ch)
    elif unet_number == 2:
        kw = dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=2, memory_efficient=True, layer_attns=(False, False, False, True),
                  layer_cross_attns=(False, False, True, True), init_conv_to_final_conv_residual=True, cond_images_channels=cond_ch)
    elif unet_number == 3:
        kw = dict(dim=128, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 6, 8), memory_efficient=True, layer_attns=False,
                  layer_cross_attns=(False, False, False, True), init_conv_to_final_conv_residual=True, cond_images_channels=cond_ch)
    else:
        return None
    kw.update(overrides)
    return Unet(**kw)


def init_imagen_ultra_res(magnification_level, unet_number, device=None, version="v_param", timesteps=(1024, 256, 256), **unet_overrides):
    objectives = {"v_param": ("noise", "v", "v"), "airs": ("v", "v", "v")}.get(version, ("noise", "noise", "noise"))
    imagen = Imagen(
        unets=(
            ultra_res_unet(magnification_level, 1, version, **unet_overrides) if unet_number == 1 else FixedNullUnet(),
            ultra_res_unet(magnification_level, 2, version, **unet_overrides) if unet_number == 2 else FixedNullUnet(lowres_cond=True),
            ultra_res_unet(magnification_level, 3, version, **unet_overrides) if unet_number == 3 else FixedNullUnet(lowres_cond=True),
        ),
        image_sizes=(64, 256, 1024), timesteps=timesteps, pred_objectives=objectives, random_crop_sizes=(None, None, 256),
        condition_on_text=False,
    )
    return imagen.to(device) if device is not None else imagen


def uncond_unet(unet_number, **overrides):
    """train_uncond.py:28-63."""
    if unet_number == 1:
        kw = dict(dim=256, dim_mults=(1, 2, 4, 8), cond_dim=512, num_resnet_blocks=3, layer_attns=(False, True, True, True),
                  layer_cross_attns=(False, True, True, True))
    elif unet_number == 2:
        kw = dict(dim=128, cond_dim=512, dim_mults=(1, 2, 4, 8), num_resnet_blocks=2, memory_efficient=True,
                  layer_attns=(False, False, False, True), layer_cross_attns=(False, False, True, True),
                  init_conv_to_final_conv_residual=True)
    else:
        kw = dict(dim=128, cond_dim=512, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 4, 4), memory_efficient=True, layer_attns=False,
                  layer_cross_attns=(False, False, False, True), init_conv_to_final_conv_residual=True)
    kw.update(overrides)
    return Unet(**kw)


def init_imagen_uncond(unet_number, device=None, timesteps=(1024, 256, 256), **unet_overrides):
    """train_uncond.py:79-93."""
    imagen = Imagen(
        condition_on_text=False,
        unets=(
            uncond_unet(1, **unet_overrides) if unet_number == 1 else FixedNullUnet(),
            uncond_unet(2, **unet_overrides) if unet_number == 2 else FixedNullUnet(lowres_cond=True),
            uncond_unet(3, **unet_overrides) if unet_number == 3 else FixedNullUnet(lowres_cond=True),
        ),
        image_sizes=(64, 256, 1024), timesteps=timesteps, pred_objectives=("noise", "noise", "noise"),
        random_crop_sizes=(None, None, 256),
    )
    return imagen.to(device) if device is not None else imagen


def cond_unet(unet_number, **overrides):
    """train.py:28-67 (segmentation-mask + clinical-vector conditioned)."""
    if unet_number == 1:
        kw = dict(dim=256, dim_mults=(1, 2, 3, 4), cond_dim=512, text_embed_dim=3, num_resnet_blocks=3,
                  layer_attns=(False, True, True, True), layer_cross_attns=(False, True, True, True), cond_images_channels=4)
    elif unet_number == 2:
        kw = dict(dim=128, cond_dim=512, dim_mults=(1, 2, 4, 8), num_resnet_blocks=2, memory_efficient=True,
                  layer_attns=(False, False, False, True), layer_cross_attns=(False, False, True, True),
                  init_conv_to_final_conv_residual=True, cond_images_channels=4)
    else:
        kw = dict(dim=128, cond_dim=512, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(2, 4, 4, 4), memory_efficient=True, layer_attns=False,
                  layer_cross_attns=(False, False, False, True), init_conv_to_final_conv_residual=True, cond_images_channels=4)
    kw.update(overrides)
    return Unet(**kw)


def init_imagen_cond(unet_number, device=None, timesteps=(1024, 256, 256), **unet_overrides):
    """train.py:83-95."""
    imagen = Imagen(
        unets=(
            cond_unet(1, **unet_overrides) if unet_number == 1 else FixedNullUnet(),
            cond_unet(2, **unet_overrides) if unet_number == 2 else FixedNullUnet(lowres_cond=True),
            cond_unet(3, **unet_overrides) if unet_number == 3 else FixedNullUnet(lowres_cond=True),
        ),
        image_sizes=(64, 256, 1024), timesteps=timesteps, pred_objectives=("noise", "v", "v"), text_embed_dim=3,
        random_crop_sizes=(None, None, 256),
    )
    return imagen.to(device) if device is not None else imagen


def randomize_zero_init_(module: nn.Module, std: float = 0.02, seed: int = 1):
    """Synthetic-weight benchmarks only: imagen-pytorch zero-initialises final_conv, so a fresh UNet predicts exactly 0."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            if p.numel() > 1 and bool((p == 0).all()):
                p.copy_((torch.randn(p.shape, generator=g) * std).to(p.device))
    return module
