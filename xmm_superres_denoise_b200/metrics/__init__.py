# mirrors the part of xmm_superres_denoise/metrics/__init__.py that is on the hot path
from .metrics import PoissonNLLLoss  # noqa: F401
from .xmm_metric_collection import (  # noqa: F401
    XMMMetricCollection,
    get_ext_metrics,
    get_in_ext_metrics,
    get_in_metrics,
    get_metrics,
)
