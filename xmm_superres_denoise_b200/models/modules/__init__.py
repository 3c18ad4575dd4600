# mirrors xmm_superres_denoise/models/modules/__init__.py:1
from .rrdb_blocks import RRDB, ResidualDenseBlock_5C, make_layer  # noqa: F401
