"""scopyon_b200 -- B200-native image formation with scopyon's public API.

Drop-in for the hot path of ecell/scopyon (``form_image``, ``generate_images``,
``sample_inputs``, ``Image``, ``DefaultConfiguration`` ...); the work runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/scopyon_b200.h``.
"""
from .base import *
from .config import *
from .image import *
from .sampling import *
from .sampling2 import *
from . import constants
from . import analysis

__all__ = [
    "EnvironSettings", "EPIFMSimulator",
    "form_image", "generate_images", "create_simulator",
    "Configuration", "DefaultConfiguration",
    "Image", "Video",
    "sample_inputs",
    "sample",
    "constants",
    "analysis",
    ]

__version__ = "0.1.0"
