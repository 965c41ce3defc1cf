"""Patch-grid sampler on the GPU: the CUDA canvas kernel inside the real scheduler, and invariance of every patch to how
patches are batched (the property that makes 1/2/4/8-GPU runs produce identical images)."""
import types

import pytest
import torch

from helpers import U1_KW, U2_KW

pytestmark = pytest.mark.gpu


def _provider(timesteps):
    from kidney_diffusion_b200.factories import FixedNullUnet, randomize_zero_init_
    from kidney_diffusion_b200 import Imagen, Unet

    def make(mag, n, device, args):
        torch.manual_seed(100 + n)
        kw1 = dict(U1_KW, cond_images_channels=3)
        unets = (Unet(**kw1) if n == 1 else FixedNullUnet(), Unet(**U2_KW) if n == 2 else FixedNullUnet(lowres_cond=True),
                 FixedNullUnet(lowres_cond=True))
        im = Imagen(unets=unets, image_sizes=(64, 256, 1024), timesteps=timesteps, pred_objectives=("noise", "v", "v"),
                    random_crop_sizes=(None, None, 256), condition_on_text=False)
        randomize_zero_init_(im)
        return im.to(device).eval()

    return make


def test_grid_stage_batch_invariance_and_borders(cuda_lib):
    from kidney_diffusion_b200 import grid

    grid._MODEL_CACHE.clear()
    grid.MODEL_PROVIDER = _provider((3, 2, 2))
    grid.CANVAS_FN = grid.default_canvas
    g = torch.Generator().manual_seed(0)
    zoomed = torch.rand(1, 3, 420, 420, generator=g)
    outs = {}
    for mb in (1, 4):
        args = types.SimpleNamespace(version="v_param", overlap=0.25, inpaint_resample=2, ignore_unet_1=False, num_gpus=1, device="cuda:0",
                                     max_batch=mb)
        cond, pos, n = grid.get_cond_images(args, zoomed, 1)
        assert n == 4 and len(pos) == 16
        o = grid.choose_orientation(pos)
        low = grid.generate_image_with_unet(1, 1, args, None, cond, pos, 0.25, o, n)
        med = grid.generate_image_with_unet(1, 2, args, low, cond, pos, 0.25, o, n)
        outs[mb] = (torch.cat(list(low)), torch.cat(list(med)))
    for a, b in zip(outs[1], outs[4]):
        assert torch.equal(a, b), "a patch must not depend on which patches share its batch"
    low, med = outs[1]
    S, ov = 256, 64
    # the overlap border of every patch equals its neighbour's facing strip (inpainted region is pasted back exactly)
    for k, (i, j) in enumerate(pos):
        if (i - 1, j) in pos:
            up = pos.index((i - 1, j))
            assert torch.equal(med[k, :, :ov, :][:, :, ov:], med[up, :, -ov:, :][:, :, ov:]) or torch.allclose(
                med[k, :, :ov, ov:], med[up, :, -ov:, ov:], atol=1e-6)
    assert float(med.min()) >= 0 and float(med.max()) <= 1
