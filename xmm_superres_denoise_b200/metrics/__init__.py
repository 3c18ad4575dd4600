# mirrors the part of xmm_superres_denoise/metrics/__init__.py that is on the hot path
from .metrics import PoissonNLLLoss  # noqa: F401
