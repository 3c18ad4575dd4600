# mirrors xmm_superres_denoise/transforms/__init__.py
from .imageupsample import ImageUpsample  # noqa: F401
from .normalize import Normalize  # noqa: F401
