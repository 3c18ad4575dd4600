# mirrors xmm_superres_denoise/models/__init__.py
from .model import Model  # noqa: F401
from .modules.generator_rrdb import GeneratorRRDB_DN, GeneratorRRDB_SR  # noqa: F401
