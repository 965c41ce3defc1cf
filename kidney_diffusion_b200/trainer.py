"""Sampling-only shim of imagen-pytorch's ``ImagenTrainer`` (reference use: sample_cond.py:26-48, sample_uncond.py:23-55)
and ``restore_parts`` (sample_ultra_res.py:63).  Training methods are out of scope.
"""
from __future__ import annotations

import warnings

import torch
from torch import nn

__version__ = "1.18.5"


def restore_parts(state_dict_target, state_dict_from):
    """Copy every tensor of ``state_dict_from`` whose name exists in the target with an equal shape."""
    for name, param in state_dict_from.items():
        if name not in state_dict_target:
            continue
        if param.size() == state_dict_target[name].size():
            state_dict_target[name].copy_(param)
        else:
            print(f"layer {name}({param.size()} different than target: {state_dict_target[name].size()}")
    return state_dict_target


class ImagenTrainer(nn.Module):
    def __init__(self, imagen=None, imagen_checkpoint_path=None, use_ema=True, fp16=False, **kwargs):
        super().__init__()
        assert imagen is not None, "pass the Imagen instance (checkpoint-path construction is not part of the sampling path)"
        self.imagen = imagen
        self.use_ema = use_ema
        self._ema_loaded = False
        self._online_backup = None

    @property
    def device(self):
        return self.imagen.device

    def load(self, path, only_model=False, strict=True, noop_if_not_exist=False):
        import os

        if noop_if_not_exist and not os.path.exists(str(path)):
            return None
        loaded_obj = torch.load(str(path), map_location="cpu")
        if loaded_obj.get("version", __version__) != __version__:
            print(f'loading saved imagen at version {loaded_obj["version"]}, but current package version is {__version__}')
        try:
            self.imagen.load_state_dict(loaded_obj["model"], strict=strict)
        except RuntimeError:
            print("Failed loading state dict. Trying partial load")
            self.imagen.load_state_dict(restore_parts(self.imagen.state_dict(), loaded_obj["model"]))
        if only_model:
            return loaded_obj
        # trainer.sample() runs the EMA copies of the unets: fold "<i>.ema_model.<key>" into the unets
        ema = loaded_obj.get("ema")
        if self.use_ema and ema is not None:
            for i, unet in enumerate(self.imagen.unets):
                prefix = f"{i}.ema_model."
                sd = {k[len(prefix):]: v for k, v in ema.items() if k.startswith(prefix)}
                if sd:
                    missing, unexpected = unet.load_state_dict(sd, strict=False)
                    if missing:
                        warnings.warn(f"EMA weights for unet {i + 1} miss {len(missing)} tensors; online weights kept for those")
            self._ema_loaded = True
        return loaded_obj

    @torch.no_grad()
    def sample(self, *args, max_batch_size=None, use_non_ema=False, **kwargs):
        kwargs.setdefault("device", self.device)
        if max_batch_size is None:
            return self.imagen.sample(*args, **kwargs)
        batch_size = kwargs.get("batch_size", 1)
        outs = []
        for b0 in range(0, batch_size, max_batch_size):
            kw = dict(kwargs)
            n = min(max_batch_size, batch_size - b0)
            kw["batch_size"] = n
            for name in ("text_embeds", "text_masks", "cond_images", "inpaint_images", "inpaint_masks", "start_image_or_video"):
                if kw.get(name) is not None:
                    kw[name] = kw[name][b0:b0 + n]
            outs.append(self.imagen.sample(*args, **kw))
        if isinstance(outs[0], list):
            return [i for o in outs for i in o]
        return torch.cat(outs, 0)
