# mirrors xmm_superres_denoise/models/__init__.py:2 (the LightningModule stays the reference's own)
from .modules.generator_rrdb import GeneratorRRDB_DN, GeneratorRRDB_SR  # noqa: F401
