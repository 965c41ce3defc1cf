"""Shared test helpers: deterministic keyed noise and oracle/product model pairs (test infrastructure only)."""
import zlib

import torch

SITE_ID = {"lowres_aug": 1, "init": 2, "inpaint": 3, "p_sample": 4, "renoise": 5}


class KeyedNoise:
    """Identical N(0,1) tensors for the oracle (CPU) and the CUDA path, keyed by (site, unet, step, r)."""

    def __init__(self, seed=1234):
        self.seed = seed

    def cpu(self, site, shape, unet=0, step=0, r=0):
        g = torch.Generator().manual_seed(zlib.crc32(f"{self.seed}/{site}/{unet}/{step}/{r}".encode()))
        return torch.randn(tuple(shape), generator=g)

    def dev(self, site, shape, device, unet=0, step=0, r=0):
        return self.cpu(site, shape, unet=unet, step=step, r=r).to(device)


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


U3_KW = dict(dim=64, dim_mults=(1, 2, 4, 8), num_resnet_blocks=(1, 2, 2, 2), memory_efficient=True, layer_attns=False,
             layer_cross_attns=(False, False, False, True), init_conv_to_final_conv_residual=True, cond_images_channels=3)
U2_KW = dict(dim=64, dim_mults=(1, 2, 4, 8), num_resnet_blocks=2, memory_efficient=True, layer_attns=(False, False, False, True),
             layer_cross_attns=(False, False, True, True), init_conv_to_final_conv_residual=True, cond_images_channels=3)
U1_KW = dict(dim=64, dim_mults=(1, 2, 3, 4), num_resnet_blocks=2, layer_attns=(False, True, True, True),
             layer_cross_attns=(False, True, True, True))


def make_pair(kw, *, lowres_cond, seed=0):
    """(oracle Unet, product Unet on cuda) with identical, fully randomised weights."""
    from kidney_diffusion_b200 import Unet
    from oracle import imagen_oracle as O

    torch.manual_seed(seed)
    ou = O.Unet(**kw, lowres_cond=lowres_cond, cond_on_text=False, text_embed_dim=None)
    O.randomize_zero_init_(ou)
    pu = Unet(**kw, lowres_cond=lowres_cond, cond_on_text=False, text_embed_dim=None)
    pu.load_state_dict(ou.state_dict())
    return ou.eval(), pu.cuda().eval()
