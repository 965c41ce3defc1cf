"""kidney_diffusion_b200 -- B200-native (sm_100a) sampling hot path of jameshball/kidney-diffusion.

Python host code (this package) mirrors the reference's imagen-pytorch API for the sampling path and calls
hand-written CUDA through the C ABI of libkidney_b200.so (include/kidney_b200.h).  No CPU fallback exists.
"""
__version__ = "0.1.0"

from .imagen import Imagen  # noqa: E402,F401
from .trainer import ImagenTrainer, restore_parts  # noqa: E402,F401
from .unet import NullUnet, Unet  # noqa: E402,F401
