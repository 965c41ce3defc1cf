"""Import shim: lets the reference's scripts (`from imagen_pytorch import Unet, ImagenTrainer, Imagen, NullUnet,
SRUnet1024, ElucidatedImagen`, train_ultra_res_v_param.py:8) resolve to the B200-native implementation."""
from kidney_diffusion_b200 import Imagen, ImagenTrainer, NullUnet, Unet  # noqa: F401
from kidney_diffusion_b200.trainer import __version__  # noqa: F401


class _NameOnly:
    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} is imported but never used by the reference; it is not built")


class SRUnet1024(_NameOnly):
    pass


class ElucidatedImagen(_NameOnly):
    pass
